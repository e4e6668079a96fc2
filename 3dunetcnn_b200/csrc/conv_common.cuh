// Shared pieces of the two implicit-GEMM convolution kernels (igemm_conv.cu: per-tap streaming tiles;
// conv_halo.cu: halo-resident tiles): argument block, TMA map bundle and the fused epilogue.
#pragma once
#include "kernels.h"
#include "ptx.cuh"
#include "tmap.h"

namespace b200 {

struct ConvMaps {
  CUtensorMap a[2][2];  // [source][hi/lo]
  CUtensorMap b[2][2];
};

struct ConvArgs {
  int N, Do, Ho, Wo, Cout;
  int tw, th, td;
  int tiles_w, tiles_h, tiles_d;
  int ntaps[2], ksz[2], kchunks[2], stride[2];
  int npass;
  int mode;
  bf16* out_hi; bf16* out_lo; int ldo;
  const bf16* res_hi; const bf16* res_lo; int ldr;
  const float* scale;
  double* stats; int stats_ld;
  const bf16* x_hi; const bf16* x_lo; int ldx;
  const float4* coef; int coef_ld;
  float slope;
  double* bstats;
};

__device__ __forceinline__ void epi_load8(const bf16* hi, const bf16* lo, long long off, float* v) {
  uint4 a = *reinterpret_cast<const uint4*>(hi + off);
  v[0] = bf16_lo_to_f(a.x); v[1] = bf16_hi_to_f(a.x); v[2] = bf16_lo_to_f(a.y); v[3] = bf16_hi_to_f(a.y);
  v[4] = bf16_lo_to_f(a.z); v[5] = bf16_hi_to_f(a.z); v[6] = bf16_lo_to_f(a.w); v[7] = bf16_hi_to_f(a.w);
  if (lo) {
    uint4 b = *reinterpret_cast<const uint4*>(lo + off);
    v[0] += bf16_lo_to_f(b.x); v[1] += bf16_hi_to_f(b.x); v[2] += bf16_lo_to_f(b.y); v[3] += bf16_hi_to_f(b.y);
    v[4] += bf16_lo_to_f(b.z); v[5] += bf16_hi_to_f(b.z); v[6] += bf16_lo_to_f(b.w); v[7] += bf16_hi_to_f(b.w);
  }
}
__device__ __forceinline__ void epi_store8(bf16* hi, bf16* lo, long long off, const float* v) {
  uint4 a;
  a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(hi + off) = a;
  if (lo) {
    uint4 b;
    b.x = pack_bf16x2(v[0] - bf16_lo_to_f(a.x), v[1] - bf16_hi_to_f(a.x));
    b.y = pack_bf16x2(v[2] - bf16_lo_to_f(a.y), v[3] - bf16_hi_to_f(a.y));
    b.z = pack_bf16x2(v[4] - bf16_lo_to_f(a.z), v[5] - bf16_hi_to_f(a.z));
    b.w = pack_bf16x2(v[6] - bf16_lo_to_f(a.w), v[7] - bf16_hi_to_f(a.w));
    *reinterpret_cast<uint4*>(lo + off) = b;
  }
}

// Transposing butterfly: every lane holds 16 column values of its own row; on return lane l holds the sum over the
// 32 rows of column ((l >> 1) & 15)  (lanes 2k and 2k+1 hold the same column).  16 shuffles instead of 80.
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 16;
    float send = up ? v[i] : v[i + 8];
    float keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 8;
    float send = up ? v[i] : v[i + 4];
    float keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 4;
    float send = up ? v[i] : v[i + 2];
    float keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    float send = up ? v[0] : v[1];
    float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}


// Fused epilogue for one 128-row accumulator tile living at TMEM address `tacc` (lane quadrant already applied by
// the caller through `lane_base`): per 16-column chunk  tcgen05.ld -> (+residual)(*scale) | GN/ReLU backward ->
// hi/lo store -> per-channel partial sums into s_stats.  `vox` = linear NDHW index of this thread's row.
template <int BN>
__device__ __forceinline__ void conv_epilogue_tile(const ConvArgs& p, uint32_t tacc, int lane_base, int lane, int n,
                                                   int n0, long long vox, bool valid, float* s_stats,
                                                   const float4* s_coef, bool want_stats) {
#pragma unroll 1
  for (int j = 0; j < BN / 16; ++j) {
    const int c0 = n0 + j * 16;
    if (c0 >= p.Cout) break;
    uint32_t r[16];
    tmem_ld16(tacc + (static_cast<uint32_t>(lane_base) << 16) + j * 16, r);
    tmem_ld_wait();
    float v[16], q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int cc = c0 + hf * 8;
      float* vv = v + hf * 8;
      float* qq = q + hf * 8;
      if (cc < p.Cout && valid) {
        if (p.mode == 0) {
          if (p.res_hi) {
            float rr[8];
            epi_load8(p.res_hi, p.res_lo, vox * p.ldr + cc, rr);
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] += rr[i];
          }
          if (p.scale) {
#pragma unroll
            for (int i = 0; i < 8; ++i) vv[i] *= __ldg(p.scale + (long long)n * p.Cout + cc + i);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) qq[i] = vv[i] * vv[i];
        } else {
          float xx[8];
          epi_load8(p.x_hi, p.x_lo, vox * p.ldx + cc, xx);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 k = s_coef[j * 16 + hf * 8 + i];
            const float z = fmaf(k.x, xx[i], k.y);
            const float dz = z > 0.f ? vv[i] : vv[i] * p.slope;
            vv[i] = dz;
            qq[i] = dz * (xx[i] - k.z) * k.w;
          }
        }
        epi_store8(p.out_hi, p.out_lo, vox * p.ldo + cc, vv);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { vv[i] = 0.f; qq[i] = 0.f; }
      }
    }
    if (want_stats) {
      const float s1 = warp_colsum16(v, lane);
      const float s2 = warp_colsum16(q, lane);
      if ((lane & 1) == 0) {
        const int col = j * 16 + ((lane >> 1) & 15);
        atomicAdd(&s_stats[col * 2 + 0], s1);
        atomicAdd(&s_stats[col * 2 + 1], s2);
      }
    }
  }
}

}  // namespace b200
