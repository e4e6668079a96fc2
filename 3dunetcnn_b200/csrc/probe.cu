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

// ---- MMA issue-rate micro-benchmark: one warp per CTA issues `reps` x `inner` tcgen05.mma (M=128, N, K=16, bf16)
// on resident (zeroed) shared-memory operands and reports clock64 cycles per MMA.  a_sbo / layout / a_step let the
// caller mimic the halo kernel's shifted, re-strided A descriptors.
__global__ void __launch_bounds__(64) k_umma_rate(int N, int layout, int a_sbo, int b_sbo, int a_step, int inner, int reps,
                                                  long long* __restrict__ out, const uint8_t* __restrict__ copy_src,
                                                  int copy_bytes, int commit_each_rep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint64_t* cbar = bar + 1;                                     // copy-warp barrier
  volatile uint32_t* done = reinterpret_cast<volatile uint32_t*>(bar + 2);
  uint64_t* sbar = bar + 3;                                     // per-"stage" commit target (never waited on)
  uint64_t* dbar = bar + 4;                                     // a barrier whose phase 0 is complete from the start
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 5);
  if (threadIdx.x == 0) *done = 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); mbar_init(cbar, 1); mbar_init(sbar, 1); mbar_init(dbar, 1); mbar_arrive(dbar); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t issue = elect_one() ? 1u : 0u;
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t hi_a = desc_hi(a_sbo, layout), hi_b = desc_hi(b_sbo, layout);
    const uint32_t a_lo = desc_lo(smem_u32(smem), 16), b_lo = desc_lo(smem_u32(smem + 96 * 1024), 16);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
      for (int i = 0; i < inner; ++i)
        umma_bf16_if(issue, tm + (i & 1) * 256, desc_from(a_lo + ((i * a_step) >> 4), hi_a), desc_from(b_lo + ((i & 3) * 2), hi_b),
                     idesc, 1u);
      // stage hand-back as a pipeline would do it: bit 0 commit, bit 1 wait on an (already complete) full barrier,
      // bit 2 tcgen05.fence::after_thread_sync
      if (commit_each_rep & 1) umma_commit_if(issue, sbar);
      if (commit_each_rep & 2) mbar_wait(dbar, 0);
      if (commit_each_rep & 4) tc_fence_after();
    }
    umma_commit_if(issue, bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x] = t1 - t0; *done = 1; }
  } else if (copy_bytes > 0 && lane == 0) {
    // operand-write pressure: bulk copies global (L2-resident) -> shared memory region [128K, 160K) until the MMAs finish
    long long copied = 0;
    uint32_t ph = 0;
    const uint32_t dst = smem_u32(smem + 128 * 1024), mb = smem_u32(cbar);
    const uint8_t* src = copy_src + (size_t)blockIdx.x * 32768;
    while (*done == 0) {
      const uint32_t q = copy_bytes / 4;   // four copies in flight per round
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(4 * q) : "memory");
#pragma unroll
      for (uint32_t c = 0; c < 4; ++c)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + c * q),
                     "l"(src + c * q), "r"(q), "r"(mb)
                     : "memory");
      mbar_wait(cbar, ph);
      ph ^= 1;
      copied += 4 * q;
    }
    out[gridDim.x + blockIdx.x] = copied;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- stage-structured issue benchmark: the MMA-warp loop of the conv kernels in isolation.  `nw` issuing warps take
// the stages round-robin; a stage = wait on an (already complete) full barrier + tcgen05 fence + one elected lane issuing
// `MPS` MMAs (M=128, N, K=16) and a commit.  Reports cycles for all stages (warp 0's clock).
template <int MPS>
__global__ void __launch_bounds__(64) k_umma_issue(int N, int stages, int nw, long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint64_t* sbar = bar + 1;
  uint64_t* dbar = bar + 2;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 3);
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, nw); mbar_init(sbar, 1); mbar_init(dbar, 1); mbar_arrive(dbar); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, *slot, 0);
  if (warp < nw) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t hi = desc_hi(1024, UMMA_SW128);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    uint32_t bs = warp, NB = 7;
    const long long t0 = clock64();
    for (int s = warp; s < stages; s += nw) {
      mbar_wait(dbar, 0);
      tc_fence_after();
      const uint32_t alo = desc_lo(a0 + (s & 7) * 1024, 16), blo = desc_lo(b0 + bs * 4096, 16);
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < MPS; ++i)
          umma_bf16(tm + (i & 1) * 256, desc_from(alo + (i >> 1) * 2 + (i & 1) * 360, hi), desc_from(blo + (i >> 1) * 2, hi), idesc, 1u);
        umma_commit(sbar);
      }
      __syncwarp();
      bs += nw;
      if (bs >= NB) bs -= NB;
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (warp == 0 && lane == 0) out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int launch_umma_rate(int N, int layout, int a_sbo, int b_sbo, int a_step, int inner, int reps, int ctas, long long* out,
                     const void* copy_src, int copy_bytes, int commit_each_rep, cudaStream_t st) {
  B200_REQUIRE(copy_bytes >= 0 && copy_bytes <= 32768 && copy_bytes % 64 == 0, E_INVALID, "umma_rate: copy_bytes=%d", copy_bytes);
  const int smem = 160 * 1024 + 64 + 1024;
  if (commit_each_rep & 24) {   // stage-structured issue benchmark: bit 3 = one issuing warp, bit 4 = two; inner = MMAs per stage
    const int nw = (commit_each_rep & 16) ? 2 : 1;
    B200_REQUIRE(inner == 4 || inner == 12, E_INVALID, "umma_rate: stage benchmark supports 4 or 12 MMAs per stage");
    if (inner == 4) {
      B200_CHECK_CUDA(cudaFuncSetAttribute(k_umma_issue<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      k_umma_issue<4><<<ctas, 64, smem, st>>>(N, reps, nw, out);
    } else {
      B200_CHECK_CUDA(cudaFuncSetAttribute(k_umma_issue<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      k_umma_issue<12><<<ctas, 64, smem, st>>>(N, reps, nw, out);
    }
    B200_CHECK_CUDA(cudaGetLastError());
    return OK;
  }
  B200_CHECK_CUDA(cudaFuncSetAttribute(k_umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_umma_rate<<<ctas, 64, smem, st>>>(N, layout, a_sbo, b_sbo, a_step, inner, reps, out, reinterpret_cast<const uint8_t*>(copy_src), copy_bytes, commit_each_rep);
  B200_CHECK_CUDA(cudaGetLastError());
  return OK;
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
