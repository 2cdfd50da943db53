// ctk_inst.cu -- explicit instantiations of the refine kernel for one (arithmetic, family) pair.
// Built once per pair with -DCTK_INST_REAL=<float|double> -DCTK_INST_FAM=<0|1|2>.
#include "ctk_kernel.cuh"

#ifndef CTK_INST_REAL
#error "define CTK_INST_REAL and CTK_INST_FAM"
#endif

namespace ctk {
#define CTK_INST(ND, ISO, SZ, EX)                                                             \
  template int launch_refine<Config<CTK_INST_REAL, ND, ISO, CTK_INST_FAM, SZ, EX> >(          \
      const BatchArgs&, cudaStream_t, char*, size_t);
#define CTK_INST_GEOM(SZ, EX)                                                                 \
  CTK_INST(2, true, SZ, EX) CTK_INST(2, false, SZ, EX) CTK_INST(3, true, SZ, EX)              \
  CTK_INST(3, false, SZ, EX)
CTK_INST_GEOM(false, false)
CTK_INST_GEOM(true, false)
#if CTK_INST_FAM != 0
CTK_INST_GEOM(false, true)
CTK_INST_GEOM(true, true)
#endif
// large-cluster instances: every derivative slot, global-memory workspace
#define CTK_INST_BIG(ND, ISO)                                                                 \
  template int launch_refine<Config<CTK_INST_REAL, ND, ISO, CTK_INST_FAM, true,               \
                                    CTK_INST_FAM != CTK_FAMILY_GAUSS, true> >(                \
      const BatchArgs&, cudaStream_t, char*, size_t);
CTK_INST_BIG(2, true) CTK_INST_BIG(2, false) CTK_INST_BIG(3, true) CTK_INST_BIG(3, false)
}  // namespace ctk
