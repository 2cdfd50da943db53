// ctk_inst.cu -- explicit instantiations of the refine kernel for one (arithmetic, family) pair.
// Built once per pair and flavour with -DCTK_INST_REAL=<float|double> -DCTK_INST_FAM=<0|1|2>
// -DCTK_INST_EXTRA=<0|1> (1: with the constraint and lowpass paths, plus the large-cluster instances).
#include "ctk_kernel.cuh"

#if !defined(CTK_INST_REAL) || !defined(CTK_INST_FAM) || !defined(CTK_INST_EXTRA)
#error "define CTK_INST_REAL, CTK_INST_FAM and CTK_INST_EXTRA"
#endif

namespace ctk {
#if CTK_INST_EXTRA
// the full flavour also carries the global-level pass (ctk_global_pass)
#define CTK_INST(ND, ISO, SZ, EX)                                                             \
  template int launch_refine<Config<CTK_INST_REAL, ND, ISO, CTK_INST_FAM, SZ, EX, false,      \
                                    true> >(const BatchArgs&, cudaStream_t, char*, size_t);   \
  template int launch_global<Config<CTK_INST_REAL, ND, ISO, CTK_INST_FAM, SZ, EX, false,      \
                                    true> >(const BatchArgs&, cudaStream_t, char*, size_t);
#else
#define CTK_INST(ND, ISO, SZ, EX)                                                             \
  template int launch_refine<Config<CTK_INST_REAL, ND, ISO, CTK_INST_FAM, SZ, EX, false,      \
                                    false> >(const BatchArgs&, cudaStream_t, char*, size_t);
#endif
#define CTK_INST_GEOM(SZ, EX)                                                                 \
  CTK_INST(2, true, SZ, EX) CTK_INST(2, false, SZ, EX) CTK_INST(3, true, SZ, EX)              \
  CTK_INST(3, false, SZ, EX)
CTK_INST_GEOM(false, false)
CTK_INST_GEOM(true, false)
#if CTK_INST_FAM != 0
CTK_INST_GEOM(false, true)
CTK_INST_GEOM(true, true)
#endif
// large-cluster instances: every derivative slot, global-memory workspace (full flavour only)
#if CTK_INST_EXTRA
#define CTK_INST_BIG(ND, ISO)                                                                 \
  template int launch_refine<Config<CTK_INST_REAL, ND, ISO, CTK_INST_FAM, true,               \
                                    CTK_INST_FAM != CTK_FAMILY_GAUSS, true> >(                \
      const BatchArgs&, cudaStream_t, char*, size_t);
CTK_INST_BIG(2, true) CTK_INST_BIG(2, false) CTK_INST_BIG(3, true) CTK_INST_BIG(3, false)
#endif
}  // namespace ctk
