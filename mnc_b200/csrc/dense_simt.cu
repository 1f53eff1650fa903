// SIMT companions of the tensor-core path: conv1_1 (Cin = 3), 2x2 ceil-mode max pooling on
// split NHWC, split-K reduction, layout converters, and an fp32-FMA implicit GEMM with the same
// contract as mnc_igemm_tc that serves as the on-device cross-check.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "mnc_b200.h"

namespace mnc {

__device__ __forceinline__ float bf2f(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ void split_f32(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ------------------------------------------------------------------ SIMT igemm
// 64 pixels x 64 output channels per CTA, 16x16 threads, 4x4 outputs per thread.
__global__ void __launch_bounds__(256)
igemm_simt_kernel(const __nv_bfloat16* __restrict__ a_hi, const __nv_bfloat16* __restrict__ a_lo,
                  int batch, int H, int W, int Cin, const __nv_bfloat16* __restrict__ w_hi,
                  const __nv_bfloat16* __restrict__ w_lo, int Cout, int taps,
                  const float* __restrict__ bias, int relu, int out_mode, void* out0, void* out1,
                  long long out_pix_stride, int out_ch_offset) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const long long M = static_cast<long long>(batch) * H * W;
  const long long m0 = static_cast<long long>(blockIdx.x) * 64;
  const int n0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const long long Ktot = static_cast<long long>(taps) * Cin;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lp = tid / 4;        // local pixel / cout row loaded by this thread
  const int lc = (tid % 4) * 4;  // 4 consecutive channels
  const long long m = m0 + lp;
  int img = 0, h = 0, w = 0;
  if (m < M) {
    img = static_cast<int>(m / (static_cast<long long>(H) * W));
    const int r = static_cast<int>(m % (static_cast<long long>(H) * W));
    h = r / W;
    w = r % W;
  }
  for (int tap = 0; tap < taps; ++tap) {
    const int dy = (taps == 9) ? tap / 3 - 1 : 0;
    const int dx = (taps == 9) ? tap % 3 - 1 : 0;
    const int hs = h + dy, ws = w + dx;
    const bool in_ok = (m < M) && hs >= 0 && hs < H && ws >= 0 && ws < W;
    const long long a_off = ((static_cast<long long>(img) * H + hs) * W + ws) * Cin;
    for (int c0 = 0; c0 < Cin; c0 += 16) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float av = 0.f;
        if (in_ok) av = bf2f(a_hi[a_off + c0 + lc + e]) + bf2f(a_lo[a_off + c0 + lc + e]);
        As[lc + e][lp] = av;
        float bv = 0.f;
        if (n0 + lp < Cout) {
          const long long b_off = static_cast<long long>(n0 + lp) * Ktot + tap * Cin + c0 + lc + e;
          bv = bf2f(w_hi[b_off]) + bf2f(w_lo[b_off]);
        }
        Bs[lc + e][lp] = bv;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = n0 + tx * 4 + j;
      if (ch >= Cout) continue;
      float v = acc[i][j];
      if (bias) v += bias[ch];
      if (relu) v = fmaxf(v, 0.f);
      const long long o = mm * out_pix_stride + out_ch_offset + ch;
      if (out_mode == 0) {
        __nv_bfloat16 hi, lo;
        split_f32(v, hi, lo);
        static_cast<__nv_bfloat16*>(out0)[o] = hi;
        static_cast<__nv_bfloat16*>(out1)[o] = lo;
      } else {
        static_cast<float*>(out0)[o] = v;
      }
    }
  }
}

// ---------------------------------------------------------------- split-K sum
// out = act(sum_s partial[s] + bias).  One thread per 4 consecutive columns (cols % 4 == 0 fast
// path: 16-byte loads of each partial plane, 8-byte stores to both bf16 planes), rows on grid.y.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long split_stride,
                     long long rows, int cols, const float* __restrict__ bias, int relu,
                     int out_mode, void* out0, void* out1, long long out_row_stride,
                     int out_ch_offset, int vec) {
  const int c4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c4 >= cols) return;
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const float* p = partial + r * cols + c4;
    if (vec) {
      for (int s = 0; s < splits; ++s) {  // fixed order: deterministic
        const float4 x = __ldcs(reinterpret_cast<const float4*>(p + s * split_stride));
        v[0] += x.x;
        v[1] += x.y;
        v[2] += x.z;
        v[3] += x.w;
      }
    } else {
      for (int s = 0; s < splits; ++s)
        for (int e = 0; e < 4 && c4 + e < cols; ++e) v[e] += p[s * split_stride + e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (bias && c4 + e < cols) v[e] += __ldg(bias + c4 + e);
      if (relu) v[e] = fmaxf(v[e], 0.f);
    }
    const long long o = r * out_row_stride + out_ch_offset + c4;
    if (out_mode == 0) {
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_f32(v[e], hi[e], lo[e]);
      if (vec) {
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out0) + o) = make_uint2(
            static_cast<uint32_t>(__bfloat16_as_ushort(hi[0])) | (static_cast<uint32_t>(__bfloat16_as_ushort(hi[1])) << 16),
            static_cast<uint32_t>(__bfloat16_as_ushort(hi[2])) | (static_cast<uint32_t>(__bfloat16_as_ushort(hi[3])) << 16));
        *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out1) + o) = make_uint2(
            static_cast<uint32_t>(__bfloat16_as_ushort(lo[0])) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo[1])) << 16),
            static_cast<uint32_t>(__bfloat16_as_ushort(lo[2])) | (static_cast<uint32_t>(__bfloat16_as_ushort(lo[3])) << 16));
      } else {
        for (int e = 0; e < 4 && c4 + e < cols; ++e) {
          static_cast<__nv_bfloat16*>(out0)[o + e] = hi[e];
          static_cast<__nv_bfloat16*>(out1)[o + e] = lo[e];
        }
      }
    } else {
      if (vec) {
        *reinterpret_cast<float4*>(static_cast<float*>(out0) + o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        for (int e = 0; e < 4 && c4 + e < cols; ++e) static_cast<float*>(out0)[o + e] = v[e];
      }
    }
  }
}

// -------------------------------------------------------------------- conv1_1
// Lane = (pixel quad, channel octet): a thread computes 4 horizontally adjacent pixels x 8 output
// channels; the 8 lanes sharing a pixel quad cover all 64 channels, so every store instruction
// writes whole 128-byte NHWC pixel rows (no partial-sector writes, which cost read-modify-write
// traffic in L2).  The 3x6x3 input window sits in registers (loads are shared through L1 by the
// 8 lanes of a quad); weights come from shared memory with 128-bit loads, 2 per 32 FMAs.
constexpr int kC11Pix = 4;

__global__ void __launch_bounds__(256)
conv1_1_kernel(const float* __restrict__ data, int batch, int H, int W,
               const float* __restrict__ weight, const float* __restrict__ bias,
               __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
  constexpr int COUT = 64;
  __shared__ __align__(16) float ws[27][COUT];
  __shared__ float bs[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
    const int co = i / 27, k = i % 27;  // Caffe weight order [co][c][ky][kx]
    ws[k][co] = weight[i];
  }
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int Wq = (W + kC11Pix - 1) / kC11Pix;
  const long long HW = static_cast<long long>(H) * W;
  const long long gq = static_cast<long long>(blockIdx.x) * (blockDim.x / 8) + threadIdx.x / 8;
  const int oc = (threadIdx.x & 7) * 8;  // first of this lane's 8 output channels
  if (gq >= static_cast<long long>(batch) * H * Wq) return;
  const int img = static_cast<int>(gq / (static_cast<long long>(H) * Wq));
  const int r = static_cast<int>(gq % (static_cast<long long>(H) * Wq));
  const int h = r / Wq, w0 = (r % Wq) * kC11Pix;
  float in[3][3][kC11Pix + 2];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int x = 0; x < kC11Pix + 2; ++x) {
        const int hs = h + ky - 1, wsx = w0 + x - 1;
        float v = 0.f;
        if (hs >= 0 && hs < H && wsx >= 0 && wsx < W)
          v = __ldg(data + (static_cast<long long>(img) * 3 + c) * HW + static_cast<long long>(hs) * W + wsx);
        in[c][ky][x] = v;
      }
  float acc[kC11Pix][8];
#pragma unroll
  for (int p = 0; p < kC11Pix; ++p)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int k = c * 9 + ky * 3 + kx;
        const float4 wa = *reinterpret_cast<const float4*>(&ws[k][oc]);
        const float4 wb = *reinterpret_cast<const float4*>(&ws[k][oc + 4]);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int p = 0; p < kC11Pix; ++p)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(in[c][ky][p + kx], wv[j], acc[p][j]);
      }
#pragma unroll
  for (int p = 0; p < kC11Pix; ++p) {
    if (w0 + p >= W) continue;
    const long long pix = (static_cast<long long>(img) * H + h) * W + w0 + p;
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = fmaxf(acc[p][2 * e] + bs[oc + 2 * e], 0.f);
      const float x1 = fmaxf(acc[p][2 * e + 1] + bs[oc + 2 * e + 1], 0.f);
      __nv_bfloat16 h0, l0, h1, l1;
      split_f32(x0, h0, l0);
      split_f32(x1, h1, l1);
      hw[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) |
              (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      lw[e] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) |
              (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    *reinterpret_cast<uint4*>(out_hi + pix * COUT + oc) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(out_lo + pix * COUT + oc) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  }
}

// ------------------------------------------------------------- 2x2 max pooling
// One thread per (pixel, 8-channel group); 16-byte loads/stores on both planes.
__global__ void maxpool2x2_split_kernel(const __nv_bfloat16* __restrict__ in_hi,
                                        const __nv_bfloat16* __restrict__ in_lo, int batch, int H,
                                        int W, int C, __nv_bfloat16* __restrict__ out_hi,
                                        __nv_bfloat16* __restrict__ out_lo) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, C8 = C / 8;
  const long long total = static_cast<long long>(batch) * Ho * Wo * C8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % C8);
    long long r = i / C8;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int img = static_cast<int>(r / Ho);
    float best[8];
    uint16_t bh[8], bl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = -3.402823466e+38f;
      bh[j] = 0;
      bl[j] = 0;
    }
    for (int dy = 0; dy < 2; ++dy) {
      const int h = ho * 2 + dy;
      if (h >= H) continue;
      for (int dx = 0; dx < 2; ++dx) {
        const int w = wo * 2 + dx;
        if (w >= W) continue;
        const long long off = ((static_cast<long long>(img) * H + h) * W + w) * C + cg * 8;
        const uint4 vh = *reinterpret_cast<const uint4*>(in_hi + off);
        const uint4 vl = *reinterpret_cast<const uint4*>(in_lo + off);
        const uint32_t hw[4] = {vh.x, vh.y, vh.z, vh.w};
        const uint32_t lw[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint16_t hb = static_cast<uint16_t>(hw[j / 2] >> ((j & 1) * 16));
          const uint16_t lb = static_cast<uint16_t>(lw[j / 2] >> ((j & 1) * 16));
          const float v = __uint_as_float(static_cast<uint32_t>(hb) << 16) +
                          __uint_as_float(static_cast<uint32_t>(lb) << 16);
          if (v > best[j]) {  // first maximum wins, as pooling_layer.cu:36
            best[j] = v;
            bh[j] = hb;
            bl[j] = lb;
          }
        }
      }
    }
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      oh[e] = static_cast<uint32_t>(bh[2 * e]) | (static_cast<uint32_t>(bh[2 * e + 1]) << 16);
      ol[e] = static_cast<uint32_t>(bl[2 * e]) | (static_cast<uint32_t>(bl[2 * e + 1]) << 16);
    }
    const long long o = ((static_cast<long long>(img) * Ho + ho) * Wo + wo) * C + cg * 8;
    *reinterpret_cast<uint4*>(out_hi + o) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(out_lo + o) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

// ---------------------------------------------------------- layout converters
__global__ void split_to_nchw_kernel(const __nv_bfloat16* __restrict__ in_hi,
                                     const __nv_bfloat16* __restrict__ in_lo, int batch, int H,
                                     int W, int C, float* __restrict__ out) {
  // tile transpose through shared memory: 32 pixels x 32 channels
  __shared__ float t[32][33];
  const long long HW = static_cast<long long>(H) * W;
  const int img = blockIdx.z;
  const long long p0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < HW && c < C) {
      const long long off = (static_cast<long long>(img) * HW + p) * C + c;
      v = bf2f(in_hi[off]) + bf2f(in_lo[off]);
    }
    t[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long p = p0 + threadIdx.x;
    if (p < HW && c < C) out[(static_cast<long long>(img) * C + c) * HW + p] = t[threadIdx.x][i];
  }
}

__global__ void nchw_to_split_kernel(const float* __restrict__ in, int batch, int C, int H, int W,
                                     __nv_bfloat16* __restrict__ out_hi,
                                     __nv_bfloat16* __restrict__ out_lo) {
  __shared__ float t[32][33];
  const long long HW = static_cast<long long>(H) * W;
  const int img = blockIdx.z;
  const long long p0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long p = p0 + threadIdx.x;
    float v = 0.f;
    if (p < HW && c < C) v = in[(static_cast<long long>(img) * C + c) * HW + p];
    t[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int c = c0 + threadIdx.x;
    if (p < HW && c < C) {
      __nv_bfloat16 hi, lo;
      split_f32(t[threadIdx.x][i], hi, lo);
      const long long off = (static_cast<long long>(img) * HW + p) * C + c;
      out_hi[off] = hi;
      out_lo[off] = lo;
    }
  }
}

__global__ void f32_to_split_kernel(const float* __restrict__ in, long long n,
                                    __nv_bfloat16* __restrict__ out_hi,
                                    __nv_bfloat16* __restrict__ out_lo) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    __nv_bfloat16 hi, lo;
    split_f32(in[i], hi, lo);
    out_hi[i] = hi;
    out_lo[i] = lo;
  }
}
__global__ void split_to_f32_kernel(const __nv_bfloat16* __restrict__ in_hi,
                                    const __nv_bfloat16* __restrict__ in_lo, long long n,
                                    float* __restrict__ out) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = bf2f(in_hi[i]) + bf2f(in_lo[i]);
}

static inline int grid_for(long long n, int block, int cap = 148 * 16) {
  long long g = (n + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}
static inline int check_launch() { return cudaGetLastError() == cudaSuccess ? MNC_OK : MNC_ERR_CUDA; }

}  // namespace mnc

using namespace mnc;

extern "C" int mnc_igemm_simt(const void* a_hi, const void* a_lo, int batch, int H, int W, int Cin,
                              const void* w_hi, const void* w_lo, int Cout, int taps,
                              const float* bias, int relu, int out_mode, void* out0, void* out1,
                              long long out_pix_stride, int out_ch_offset, void* stream) {
  if (Cin % 16 != 0 || (taps != 1 && taps != 9)) return MNC_ERR_ARG;
  const long long M = static_cast<long long>(batch) * H * W;
  dim3 grid(static_cast<unsigned>((M + 63) / 64), static_cast<unsigned>((Cout + 63) / 64));
  igemm_simt_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a_hi), static_cast<const __nv_bfloat16*>(a_lo), batch, H,
      W, Cin, static_cast<const __nv_bfloat16*>(w_hi), static_cast<const __nv_bfloat16*>(w_lo),
      Cout, taps, bias, relu, out_mode, out0, out1, out_pix_stride, out_ch_offset);
  return check_launch();
}

extern "C" int mnc_splitk_reduce(const float* partial, int splits, long long split_stride,
                                 long long rows, int cols, const float* bias, int relu,
                                 int out_mode, void* out0, void* out1, long long out_row_stride,
                                 int out_ch_offset, void* stream) {
  if (rows <= 0 || cols <= 0) return MNC_OK;
  const int vec = (cols % 4 == 0) && (split_stride % 4 == 0) && (out_row_stride % 4 == 0) &&
                  (out_ch_offset % 4 == 0) && (reinterpret_cast<uintptr_t>(partial) % 16 == 0) &&
                  (reinterpret_cast<uintptr_t>(out0) % 16 == 0) &&
                  (out_mode == 1 || reinterpret_cast<uintptr_t>(out1) % 8 == 0);
  const int tx = (cols + 3) / 4;
  const int block = tx >= 256 ? 256 : (tx >= 128 ? 128 : (tx >= 64 ? 64 : 32));
  dim3 grid((tx + block - 1) / block, static_cast<unsigned>(rows < 32768 ? rows : 32768));
  splitk_reduce_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      partial, splits, split_stride, rows, cols, bias, relu, out_mode, out0, out1, out_row_stride,
      out_ch_offset, vec);
  return check_launch();
}

extern "C" int mnc_conv1_1(const float* data_nchw, int batch, int H, int W, const float* weight,
                           const float* bias, int Cout, void* out_hi, void* out_lo, void* stream) {
  if (Cout != 64) return MNC_ERR_ARG;
  const long long quads = static_cast<long long>(batch) * H * ((W + kC11Pix - 1) / kC11Pix);
  conv1_1_kernel<<<static_cast<unsigned>((quads + 31) / 32), 256, 0,
                   static_cast<cudaStream_t>(stream)>>>(
      data_nchw, batch, H, W, weight, bias, static_cast<__nv_bfloat16*>(out_hi),
      static_cast<__nv_bfloat16*>(out_lo));
  return check_launch();
}

extern "C" int mnc_maxpool2x2_split(const void* in_hi, const void* in_lo, int batch, int H, int W,
                                    int C, void* out_hi, void* out_lo, void* stream) {
  if (C % 8 != 0) return MNC_ERR_ARG;
  const long long total = static_cast<long long>(batch) * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  maxpool2x2_split_kernel<<<grid_for(total, 256, 148 * 32), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_hi), static_cast<const __nv_bfloat16*>(in_lo), batch, H,
      W, C, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo));
  return check_launch();
}

extern "C" int mnc_split_to_nchw(const void* in_hi, const void* in_lo, int batch, int H, int W,
                                 int C, float* out_nchw, void* stream) {
  const long long HW = static_cast<long long>(H) * W;
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), static_cast<unsigned>((C + 31) / 32), batch);
  split_to_nchw_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_hi), static_cast<const __nv_bfloat16*>(in_lo), batch, H,
      W, C, out_nchw);
  return check_launch();
}

extern "C" int mnc_nchw_to_split(const float* in_nchw, int batch, int C, int H, int W,
                                 void* out_hi, void* out_lo, void* stream) {
  const long long HW = static_cast<long long>(H) * W;
  dim3 grid(static_cast<unsigned>((HW + 31) / 32), static_cast<unsigned>((C + 31) / 32), batch);
  nchw_to_split_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
      in_nchw, batch, C, H, W, static_cast<__nv_bfloat16*>(out_hi),
      static_cast<__nv_bfloat16*>(out_lo));
  return check_launch();
}

extern "C" int mnc_f32_to_split(const float* in, long long n, void* out_hi, void* out_lo,
                                void* stream) {
  f32_to_split_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, n, static_cast<__nv_bfloat16*>(out_hi), static_cast<__nv_bfloat16*>(out_lo));
  return check_launch();
}
extern "C" int mnc_split_to_f32(const void* in_hi, const void* in_lo, long long n, float* out,
                                void* stream) {
  split_to_f32_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_hi), static_cast<const __nv_bfloat16*>(in_lo), n, out);
  return check_launch();
}
