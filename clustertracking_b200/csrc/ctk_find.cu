// ctk_find.cu -- local maxima of a batch of frames (the step in front of the refinement, SURVEY
// section 8f rank 3): the image part of clustertracking/find.py:grey_dilation (find.py:219-270)
//
//   threshold = np.percentile(image[image != 0], percentile)               find.py:210-216
//   dilation  = scipy.ndimage.grey_dilation(image, size, mode='constant')  find.py:255-259
//   maxima    = (image == dilation) & (image > threshold)                  find.py:260
//   positions in C order, minus those closer than `margin` to an edge      find.py:264-270
//
// on the device.  All kernels are HBM-bound streaming passes over integer frames:
//   histogram   one pass, shared-memory bins (uint8) or global bins (uint16)
//   threshold   numpy's 'linear' percentile from the histogram, one thread per frame
//   max filter  one separable pass per axis (flat box, zero beyond the frame edge)
//   flag/count  per 1024-pixel segment, then a per-frame scan, then an ordered write
// The pair-wise `drop_close` step (find.py:166-206) stays on the host (clustertracking_b200/find.py).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "ctk.h"

namespace {

constexpr int SEG = 1024;            // pixels per flag/count segment (= threads per block)

template <class T> struct Bins;
template <> struct Bins<uint8_t> { static constexpr int N = 256; };
template <> struct Bins<uint16_t> { static constexpr int N = 65536; };

// ---- histogram of one frame per blockIdx.y ------------------------------------------------------
// uint8: 16-byte loads and 8 replicated sets of shared-memory bins (noise images put most pixels
// into a handful of values; replication spreads the atomic conflicts); uint16: global bins
template <class T>
__global__ void __launch_bounds__(256) hist_kernel(const void* const* frames, int64_t n_pixels,
                                                   unsigned int* hist) {
  constexpr int NB = Bins<T>::N;
  const T* src = reinterpret_cast<const T*>(frames[blockIdx.y]);
  unsigned int* out = hist + (size_t) blockIdx.y * NB;
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  const int64_t tid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (NB == 256) {
    __shared__ unsigned int local[8][256];
    for (int k = threadIdx.x; k < 8 * 256; k += blockDim.x) (&local[0][0])[k] = 0;
    __syncthreads();
    unsigned int* mine = local[threadIdx.x & 7];
    const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
    const int64_t n_vec = aligned ? n_pixels / 16 : 0;
    const uint4* v = reinterpret_cast<const uint4*>(src);
    for (int64_t i = tid; i < n_vec; i += stride) {
      const uint4 raw = __ldg(v + i);
      const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        atomicAdd(&mine[w[k] & 255u], 1u);
        atomicAdd(&mine[(w[k] >> 8) & 255u], 1u);
        atomicAdd(&mine[(w[k] >> 16) & 255u], 1u);
        atomicAdd(&mine[w[k] >> 24], 1u);
      }
    }
    for (int64_t i = n_vec * 16 + tid; i < n_pixels; i += stride) atomicAdd(&mine[(unsigned) src[i]], 1u);
    __syncthreads();
    unsigned int total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) total += local[k][threadIdx.x];
    if (total) atomicAdd(&out[threadIdx.x], total);
  } else {
    for (int64_t i = tid; i < n_pixels; i += stride) atomicAdd(&out[(unsigned) src[i]], 1u);
  }
}

// ---- np.percentile(not_black, q) (numpy 'linear' method) from the histogram ---------------------
// virtual index h = (n - 1) * (q / 100); a, b = sorted[floor(h)], sorted[floor(h) + 1];
// t = h - floor(h); a + (b - a) t, or b - (b - a)(1 - t) when t >= 0.5.  No FMA contraction.
__global__ void threshold_kernel(const unsigned int* hist, int n_bins, int n_frames,
                                 double percentile, double* threshold) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const unsigned int* h = hist + (size_t) f * n_bins;
  long long n = 0;
  for (int v = 1; v < n_bins; ++v) n += h[v];
  if (n == 0) { threshold[f] = NAN; return; }
  const double q = __ddiv_rn(percentile, 100.0);
  const double virt = __dmul_rn((double) (n - 1), q);
  long long prev, next;
  if (virt >= (double) (n - 1)) { prev = next = n - 1; }
  else if (virt < 0.) { prev = next = 0; }
  else { prev = (long long) floor(virt); next = prev + 1; }
  const double t = __dsub_rn(virt, floor(virt));
  double a = 0., b = 0.;
  long long seen = 0;
  bool have_a = false;
  for (int v = 1; v < n_bins; ++v) {
    seen += h[v];
    if (!have_a && seen > prev) { a = (double) v; have_a = true; }
    if (seen > next) { b = (double) v; break; }
  }
  const double diff = __dsub_rn(b, a);
  double out = __dadd_rn(a, __dmul_rn(diff, t));
  if (t >= 0.5) out = __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, t)));
  threshold[f] = out;
}

// ---- flat max filter along one axis: out[i] = max(in[i - before .. i + after]), zero outside ------
// the frames are [n_frames] separate arrays of shape (d0, d1, d2) (2D: d0 = 1); `src_ptrs` is used
// for the first pass (caller's frames), `src_flat` for the following ones (workspace)
template <class T>
__global__ void __launch_bounds__(256) maxfilter_kernel(const void* const* src_ptrs, const T* src_flat,
                                                        T* dst, int d0, int d1, int d2, int axis,
                                                        int before, int after) {
  const int64_t n_pixels = (int64_t) d0 * d1 * d2;
  const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pixels) return;
  const T* src = src_ptrs ? reinterpret_cast<const T*>(src_ptrs[blockIdx.y])
                          : src_flat + (size_t) blockIdx.y * n_pixels;
  const int x = (int) (i % d2), y = (int) ((i / d2) % d1), z = (int) (i / ((int64_t) d1 * d2));
  const int pos = axis == 0 ? z : (axis == 1 ? y : x);
  const int len = axis == 0 ? d0 : (axis == 1 ? d1 : d2);
  const int64_t step = axis == 0 ? (int64_t) d1 * d2 : (axis == 1 ? d2 : 1);
  int lo = pos - before, hi = pos + after;
  T best = (lo < 0 || hi >= len) ? (T) 0 : src[i];          // the zero padding takes part in the max
  if (lo < 0) lo = 0;
  if (hi >= len) hi = len - 1;
  const T* p = src + i + (int64_t) (lo - pos) * step;
  for (int k = lo; k <= hi; ++k, p += step) { const T v = __ldg(p); best = v > best ? v : best; }
  dst[(size_t) blockIdx.y * n_pixels + i] = best;
}

// uint8 frames whose rows are a multiple of 4 pixels: one thread produces 4 neighbouring outputs
// from 32-bit loads with the per-byte SIMD maximum (__vmaxu4); along x the window slides through
// funnel shifts of neighbouring words
__global__ void __launch_bounds__(256) maxfilter4_kernel(const void* const* src_ptrs,
                                                         const uint8_t* src_flat, uint8_t* dst, int d0,
                                                         int d1, int d2, int axis, int before,
                                                         int after) {
  const int w2 = d2 >> 2;                                   // words per row
  const int n_words = d0 * d1 * w2;                         // < 2^29 (checked by the caller)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // 32-bit index arithmetic throughout
  if (i >= n_words) return;
  const size_t n_pixels = (size_t) n_words * 4;
  const uint8_t* src8 = src_ptrs ? reinterpret_cast<const uint8_t*>(src_ptrs[blockIdx.y])
                                 : src_flat + (size_t) blockIdx.y * n_pixels;
  const unsigned* src = reinterpret_cast<const unsigned*>(src8);
  const int row_id = i / w2;                                // z * d1 + y
  const int xw = i - row_id * w2, z = row_id / d1, y = row_id - z * d1;
  unsigned best = 0u;                                       // zero padding / smallest value
  if (axis == 2) {
    const unsigned* row = src + (i - xw);
    for (int s = -before; s <= after; ++s) {
      const int k = 4 * xw + s;                             // first byte of the shifted word
      const int wi = k >> 2, sh = (k & 3) * 8;
      const unsigned lo = (wi >= 0 && wi < w2) ? __ldg(row + wi) : 0u;
      const unsigned hi = (sh && wi + 1 >= 0 && wi + 1 < w2) ? __ldg(row + wi + 1) : 0u;
      best = __vmaxu4(best, __funnelshift_r(lo, hi, sh));
    }
  } else {
    const int pos = axis == 0 ? z : y, len = axis == 0 ? d0 : d1;
    const int step = axis == 0 ? d1 * w2 : w2;
    int lo = pos - before, hi = pos + after;
    if (lo < 0) lo = 0;
    if (hi >= len) hi = len - 1;
    const unsigned* p = src + i + (lo - pos) * step;
    for (int k = lo; k <= hi; ++k, p += step) best = __vmaxu4(best, __ldg(p));
  }
  reinterpret_cast<unsigned*>(dst + (size_t) blockIdx.y * n_pixels)[i] = best;
}

template <class T>
__device__ __forceinline__ bool is_maximum(const T* img, const T* dil, int64_t i, double thr, int d0,
                                           int d1, int d2, int m0, int m1, int m2) {
  const T v = img[i];
  if (!(v == dil[i]) || !((double) v > thr)) return false;   // NaN threshold: nothing passes
  const int x = (int) (i % d2), y = (int) ((i / d2) % d1), z = (int) (i / ((int64_t) d1 * d2));
  // near_edge = (pos < margin) | (pos > shape - margin - 1)                  find.py:266-267
  return !(x < m2 || x > d2 - m2 - 1 || y < m1 || y > d1 - m1 - 1 || z < m0 || z > d0 - m0 - 1);
}

// ---- count the maxima of every 1024-pixel segment ------------------------------------------------
template <class T>
__global__ void __launch_bounds__(SEG) count_kernel(const void* const* frames, const T* dil,
                                                    const double* threshold, int d0, int d1, int d2,
                                                    int m0, int m1, int m2, int* seg_count) {
  const int64_t n_pixels = (int64_t) d0 * d1 * d2;
  const int64_t i = (int64_t) blockIdx.x * SEG + threadIdx.x;
  const T* img = reinterpret_cast<const T*>(frames[blockIdx.y]);
  const bool flag = i < n_pixels && is_maximum(img, dil + (size_t) blockIdx.y * n_pixels, i,
                                               threshold[blockIdx.y], d0, d1, d2, m0, m1, m2);
  const int total = __syncthreads_count(flag);
  if (threadIdx.x == 0) seg_count[(size_t) blockIdx.y * gridDim.x + blockIdx.x] = total;
}

// ---- packed uint8 variants: one thread tests 4 neighbouring pixels (one 32-bit word) ---------------
// 4-bit mask of the maxima among pixels 4 i .. 4 i + 3 (a row is a whole number of words)
__device__ __forceinline__ unsigned maxima4(const unsigned* img, const unsigned* dil, int i,
                                            double thr, int d0, int d1, int d2, int m0, int m1, int m2) {
  if (!(thr == thr)) return 0u;                              // all-black frame
  const unsigned v = __ldg(img + i), d = __ldg(dil + i);
  unsigned eq = __vcmpeq4(v, d);                             // 0xff per equal byte
  if (!eq) return 0u;
  const int w2 = d2 >> 2;
  const int row_id = i / w2;
  const int xw = i - row_id * w2, z = row_id / d1, y = row_id - z * d1;
  if (y < m1 || y > d1 - m1 - 1 || z < m0 || z > d0 - m0 - 1) return 0u;
  const int cut = (int) floor(thr);                          // integer v: v > thr  <=>  v > floor(thr)
  unsigned mask = 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = 4 * xw + k;
    const int val = (int) ((v >> (8 * k)) & 255u);
    if (((eq >> (8 * k)) & 1u) && val > cut && !(x < m2 || x > d2 - m2 - 1)) mask |= 1u << k;
  }
  return mask;
}

__global__ void __launch_bounds__(SEG / 4) count4_kernel(const void* const* frames, const uint8_t* dil,
                                                         const double* threshold, int d0, int d1, int d2,
                                                         int m0, int m1, int m2, int* seg_count) {
  const int64_t n_pixels = (int64_t) d0 * d1 * d2;
  const int n_words = (int) (n_pixels >> 2);
  const int i = blockIdx.x * (SEG / 4) + threadIdx.x;
  unsigned mask = 0u;
  if (i < n_words)
    mask = maxima4(reinterpret_cast<const unsigned*>(frames[blockIdx.y]),
                   reinterpret_cast<const unsigned*>(dil + (size_t) blockIdx.y * n_pixels), i,
                   threshold[blockIdx.y], d0, d1, d2, m0, m1, m2);
  int n = __popc(mask);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) n += __shfl_xor_sync(0xffffffffu, n, m);
  __shared__ int part[SEG / 128];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
#pragma unroll
    for (int w = 0; w < SEG / 128; ++w) total += part[w];
    seg_count[(size_t) blockIdx.y * gridDim.x + blockIdx.x] = total;
  }
}

__global__ void __launch_bounds__(SEG / 4) write4_kernel(const void* const* frames, const uint8_t* dil,
                                                         const double* threshold, int d0, int d1, int d2,
                                                         int m0, int m1, int m2, const int* seg_offset,
                                                         int ndim, int capacity, int32_t* coords,
                                                         int32_t* values) {
  const int64_t n_pixels = (int64_t) d0 * d1 * d2;
  const int n_words = (int) (n_pixels >> 2);
  const int i = blockIdx.x * (SEG / 4) + threadIdx.x;
  const unsigned* img = reinterpret_cast<const unsigned*>(frames[blockIdx.y]);
  unsigned mask = 0u;
  if (i < n_words)
    mask = maxima4(img, reinterpret_cast<const unsigned*>(dil + (size_t) blockIdx.y * n_pixels), i,
                   threshold[blockIdx.y], d0, d1, d2, m0, m1, m2);
  const int mine = __popc(mask);
  int incl = mine;                                           // inclusive scan over the warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, incl, m);
    if (lane >= m) incl += o;
  }
  __shared__ int warp_total[SEG / 128];
  if (lane == 31) warp_total[warp] = incl;
  __syncthreads();
  if (!mask) return;
  int rank = seg_offset[(size_t) blockIdx.y * gridDim.x + blockIdx.x] + incl - mine;
  for (int w = 0; w < warp; ++w) rank += warp_total[w];
  const int w2 = d2 >> 2;
  const int row_id = i / w2;
  const int xw = i - row_id * w2, z = row_id / d1, y = row_id - z * d1;
  const unsigned v = __ldg(img + i);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!((mask >> k) & 1u)) continue;
    if (rank < capacity) {
      int32_t* c = coords + ((size_t) blockIdx.y * capacity + rank) * ndim;
      if (ndim == 3) { c[0] = z; c[1] = y; c[2] = 4 * xw + k; } else { c[0] = y; c[1] = 4 * xw + k; }
      values[(size_t) blockIdx.y * capacity + rank] = (int32_t) ((v >> (8 * k)) & 255u);
    }
    ++rank;
  }
}

// ---- per-frame exclusive scan of the segment counts (one block per frame) ------------------------
__global__ void __launch_bounds__(1024) scan_kernel(int* seg_count, int n_seg, int* frame_count) {
  int* c = seg_count + (size_t) blockIdx.x * n_seg;
  __shared__ int carry;
  __shared__ int part[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_seg; base += 1024) {
    const int k = base + threadIdx.x;
    const int v = k < n_seg ? c[k] : 0;
    int incl = v;                                            // inclusive scan inside the warp
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, m);
      if ((threadIdx.x & 31) >= m) incl += o;
    }
    if ((threadIdx.x & 31) == 31) part[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int p = part[threadIdx.x];
#pragma unroll
      for (int m = 1; m < 32; m <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, p, m);
        if (threadIdx.x >= m) p += o;
      }
      part[threadIdx.x] = p;
    }
    __syncthreads();
    const int before = carry + (threadIdx.x >= 32 ? part[(threadIdx.x >> 5) - 1] : 0) + incl - v;
    if (k < n_seg) c[k] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) frame_count[blockIdx.x] = carry;
}

// ---- ordered write of the maxima: positions (C order) and pixel values ---------------------------
template <class T>
__global__ void __launch_bounds__(SEG) write_kernel(const void* const* frames, const T* dil,
                                                    const double* threshold, int d0, int d1, int d2,
                                                    int m0, int m1, int m2, const int* seg_offset,
                                                    int ndim, int capacity, int32_t* coords,
                                                    int32_t* values) {
  const int64_t n_pixels = (int64_t) d0 * d1 * d2;
  const int64_t i = (int64_t) blockIdx.x * SEG + threadIdx.x;
  const T* img = reinterpret_cast<const T*>(frames[blockIdx.y]);
  const bool flag = i < n_pixels && is_maximum(img, dil + (size_t) blockIdx.y * n_pixels, i,
                                               threshold[blockIdx.y], d0, d1, d2, m0, m1, m2);
  __shared__ int warp_total[32];
  const unsigned ball = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_total[warp] = __popc(ball);
  __syncthreads();
  if (!flag) return;
  int rank = seg_offset[(size_t) blockIdx.y * gridDim.x + blockIdx.x] + __popc(ball & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) rank += warp_total[w];
  if (rank >= capacity) return;
  const int x = (int) (i % d2), y = (int) ((i / d2) % d1), z = (int) (i / ((int64_t) d1 * d2));
  int32_t* c = coords + ((size_t) blockIdx.y * capacity + rank) * ndim;
  if (ndim == 3) { c[0] = z; c[1] = y; c[2] = x; } else { c[0] = y; c[1] = x; }
  values[(size_t) blockIdx.y * capacity + rank] = (int32_t) img[i];
}

thread_local char g_find_error[256] = "";

template <class T>
int find_maxima(const void* const* d_frames, int n_frames, int d0, int d1, int d2, int ndim,
                const int* size, double percentile, const int* margin, int capacity,
                int32_t* d_coords, int32_t* d_values, int32_t* d_count, double* d_threshold,
                char* ws, cudaStream_t st, bool frames_aligned) {
  constexpr int NB = Bins<T>::N;
  const int64_t n_pixels = (int64_t) d0 * d1 * d2;
  const int n_seg = (int) ((n_pixels + SEG - 1) / SEG);
  // workspace layout: histogram | two frame-sized buffers per frame | segment counts
  size_t off = 0;
  unsigned int* hist = reinterpret_cast<unsigned int*>(ws + off);
  off += ((size_t) n_frames * NB * 4 + 255) / 256 * 256;
  T* buf_a = reinterpret_cast<T*>(ws + off);
  off += ((size_t) n_frames * n_pixels * sizeof(T) + 255) / 256 * 256;
  T* buf_b = reinterpret_cast<T*>(ws + off);
  off += ((size_t) n_frames * n_pixels * sizeof(T) + 255) / 256 * 256;
  int* seg = reinterpret_cast<int*>(ws + off);

  cudaMemsetAsync(hist, 0, (size_t) n_frames * NB * 4, st);
  int hb = (int) ((n_pixels + 256 * 16 - 1) / (256 * 16));
  if (hb < 1) hb = 1;
  if (hb > 1024) hb = 1024;
  hist_kernel<T><<<dim3(hb, n_frames), 256, 0, st>>>(d_frames, n_pixels, hist);
  threshold_kernel<<<(n_frames + 31) / 32, 32, 0, st>>>(hist, NB, n_frames, percentile, d_threshold);
  const dim3 grid((unsigned) ((n_pixels + 255) / 256), n_frames);
  const int dims[3] = {d0, d1, d2};
  // packed path: uint8, rows of whole 32-bit words, every frame 4-byte aligned (checked by the caller)
  const bool packed = sizeof(T) == 1 && (d2 & 3) == 0 && frames_aligned;
  const T* cur = nullptr;
  T* nxt = buf_a;
  bool first = true;
  for (int k = 0; k < ndim; ++k) {
    const int axis = 3 - ndim + k;
    if (dims[axis] <= 0) continue;
    const int s = size[k];
    if (s <= 1 && !first) continue;                       // a box of one pixel: identity
    const int before = s >= 1 ? (s - 1) / 2 : 0, after = s >= 1 ? s / 2 : 0;
    if (packed) {
      const dim3 grid4((unsigned) ((n_pixels / 4 + 255) / 256), n_frames);
      maxfilter4_kernel<<<grid4, 256, 0, st>>>(first ? d_frames : nullptr,
                                               reinterpret_cast<const uint8_t*>(cur),
                                               reinterpret_cast<uint8_t*>(nxt), d0, d1, d2, axis,
                                               before, after);
    } else {
      maxfilter_kernel<T><<<grid, 256, 0, st>>>(first ? d_frames : nullptr, cur, nxt, d0, d1, d2, axis,
                                                before, after);
    }
    cur = nxt;
    nxt = (nxt == buf_a) ? buf_b : buf_a;
    first = false;
  }
  const int m0 = ndim == 3 ? margin[0] : 0, m1 = margin[ndim - 2], m2 = margin[ndim - 1];
  const dim3 sgrid(n_seg, n_frames);
  if (packed) {
    const uint8_t* dil8 = reinterpret_cast<const uint8_t*>(cur);
    count4_kernel<<<sgrid, SEG / 4, 0, st>>>(d_frames, dil8, d_threshold, d0, d1, d2, m0, m1, m2, seg);
    scan_kernel<<<n_frames, 1024, 0, st>>>(seg, n_seg, d_count);
    write4_kernel<<<sgrid, SEG / 4, 0, st>>>(d_frames, dil8, d_threshold, d0, d1, d2, m0, m1, m2, seg,
                                             ndim, capacity, d_coords, d_values);
  } else {
    count_kernel<T><<<sgrid, SEG, 0, st>>>(d_frames, cur, d_threshold, d0, d1, d2, m0, m1, m2, seg);
    scan_kernel<<<n_frames, 1024, 0, st>>>(seg, n_seg, d_count);
    write_kernel<T><<<sgrid, SEG, 0, st>>>(d_frames, cur, d_threshold, d0, d1, d2, m0, m1, m2, seg, ndim,
                                           capacity, d_coords, d_values);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_find_error, sizeof(g_find_error), "ctk_find_maxima: %s", cudaGetErrorString(e));
    return CTK_E_CUDA;
  }
  return 0;
}

}  // namespace

extern "C" {

const char* ctk_find_last_error(void) { return g_find_error; }

size_t ctk_find_workspace_bytes(int32_t n_frames, int64_t n_pixels, int32_t pixel_dtype) {
  if (n_frames <= 0 || n_pixels <= 0) return 0;
  const size_t px = pixel_dtype == CTK_PIXEL_U8 ? 1 : (pixel_dtype == CTK_PIXEL_U16 ? 2 : 0);
  if (!px) return 0;
  const size_t bins = px == 1 ? 256 : 65536;
  const size_t n_seg = (size_t) ((n_pixels + SEG - 1) / SEG);
  size_t total = ((size_t) n_frames * bins * 4 + 255) / 256 * 256;
  total += 2 * (((size_t) n_frames * (size_t) n_pixels * px + 255) / 256 * 256);
  total += (size_t) n_frames * n_seg * 4 + 256;
  return total;
}

int ctk_find_maxima(const void* const* d_frames, int32_t n_frames, const int64_t* frame_shape,
                    int32_t ndim, int32_t pixel_dtype, const int32_t* size, double percentile,
                    const int32_t* margin, int32_t capacity, int32_t* d_coords_out,
                    int32_t* d_values_out, int32_t* d_count_out, double* d_threshold_out,
                    void* d_workspace, int32_t frames_aligned, void* stream) {
  if (n_frames <= 0) return 0;
  if (!d_frames || !frame_shape || (ndim != 2 && ndim != 3) || !size || !margin || capacity < 1 ||
      !d_coords_out || !d_values_out || !d_count_out || !d_threshold_out || !d_workspace) {
    snprintf(g_find_error, sizeof(g_find_error), "ctk_find_maxima: bad argument");
    return CTK_E_INVALID;
  }
  for (int k = 0; k < ndim; ++k)
    if (frame_shape[k] < 1 || frame_shape[k] > (1 << 24) || size[k] < 0 || size[k] > 255 || margin[k] < 0) {
      snprintf(g_find_error, sizeof(g_find_error), "ctk_find_maxima: bad shape, size or margin");
      return CTK_E_INVALID;
    }
  const int d0 = ndim == 3 ? (int) frame_shape[0] : 1;
  const int d1 = (int) frame_shape[ndim - 2], d2 = (int) frame_shape[ndim - 1];
  if ((int64_t) d0 * d1 * d2 > ((int64_t) 1 << 31) - SEG) {
    snprintf(g_find_error, sizeof(g_find_error), "ctk_find_maxima: frame too large");
    return CTK_E_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(d_workspace);
  if (pixel_dtype == CTK_PIXEL_U8)
    return find_maxima<uint8_t>(d_frames, n_frames, d0, d1, d2, ndim, size, percentile, margin,
                                capacity, d_coords_out, d_values_out, d_count_out, d_threshold_out,
                                ws, st, frames_aligned != 0);
  if (pixel_dtype == CTK_PIXEL_U16)
    return find_maxima<uint16_t>(d_frames, n_frames, d0, d1, d2, ndim, size, percentile, margin,
                                 capacity, d_coords_out, d_values_out, d_count_out, d_threshold_out,
                                 ws, st, false);
  snprintf(g_find_error, sizeof(g_find_error), "ctk_find_maxima: integer frames only (uint8, uint16)");
  return CTK_E_UNSUPPORTED;
}

}  // extern "C"
