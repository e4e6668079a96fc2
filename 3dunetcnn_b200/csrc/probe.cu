// Descriptor-semantics probe: issues single-CTA tcgen05.mma tiles with shifted / re-strided shared-memory descriptors
// so the host can read back which A element the tensor core fetched for every (row, k).  Used to decide whether a
// halo-resident activation box can serve all 27 taps through descriptor offsets (see DESIGN.md "next").
#include "kernels.h"
#include "ptx.cuh"

namespace b200 {

__global__ void __launch_bounds__(128) k_umma_probe(int layout_mode, int start_off, int sbo, int lbo, int base_offset,
                                                    int encoding, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // 64 KB region for A (up to 512 rows)
  uint8_t* sB = smem + 65536;         // 8 KB: B [64 rows][64 k] SW128 K-major, identity
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  // ---- fill A
  for (int i = tid; i < 512 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    const float val = encoding == 0 ? (float)(r & 255) : (float)k;
    uint32_t off;
    if (layout_mode == 0) {
      off = r * 128 + k * 2;
      off ^= ((off >> 7) & 7) << 4;   // 128B swizzle on absolute (1024-aligned base) offsets == what TMA writes
    } else {
      off = (k / 8) * 8192 + r * 16 + (k % 8) * 2;  // interleaved: [k-chunk][row][8 elems], chunk pitch 8192 B
    }
    *reinterpret_cast<bf16*>(sA + off) = __float2bfloat16_rn(val);
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    uint32_t off = r * 128 + k * 2;
    off ^= ((off >> 7) & 7) << 4;
    *reinterpret_cast<bf16*>(sB + off) = __float2bfloat16_rn(r == k ? 1.f : 0.f);
  }
  fence_proxy_async();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t a0 = smem_u32(sA) + start_off, b0 = smem_u32(sB);
    for (int k = 0; k < 4; ++k) {
      uint64_t da, db;
      if (layout_mode == 0) da = make_smem_desc(a0 + k * 32, 16, sbo, UMMA_SW128, base_offset);
      else da = make_smem_desc(a0 + k * 2 * lbo, lbo, sbo, UMMA_SW_NONE, 0);
      db = make_smem_desc(b0 + k * 32, 16, 1024, UMMA_SW128, 0);
      umma_bf16(tmem, da, db, idesc, k > 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int j = 0; j < 4; ++j) {
    uint32_t r[16];
    tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + j * 16, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + j * 16 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int launch_umma_probe(const int* tests, int ntests, float* out, cudaStream_t st) {
  const int smem = 65536 + 8192 + 64 + 1024;
  B200_CHECK_CUDA(cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int t = 0; t < ntests; ++t) {
    const int* p = tests + t * 5;
    for (int enc = 0; enc < 2; ++enc) {
      k_umma_probe<<<1, 128, smem, st>>>(p[0], p[1], p[2], p[3], p[4], enc, out + ((long long)t * 2 + enc) * 128 * 64);
      B200_CHECK_CUDA(cudaGetLastError());
    }
  }
  return OK;
}

}  // namespace b200
