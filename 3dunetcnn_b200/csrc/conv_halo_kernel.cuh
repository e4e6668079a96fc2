// Kernel template of the halo-resident convolution (see conv_halo.cu for the design notes); included by the translation units
// that instantiate its epilogue variants (conv_halo.cu: generic, conv_halo_ev1.cu / conv_halo_ev2.cu: lean).
#pragma once
#include <cstdlib>
#include "conv_common.cuh"

namespace b200 {

// NI = number of MMA-issuing warps (1, or 2 = experimental: each issuer accumulates its stages into its OWN accumulator
// set and the epilogue adds the two; BN <= 32 only, enabled with B200UNET_HALO_ISSUERS=2)
template <int KC, int BN, int TD, int NI = 1, int KW = 1>
struct HaloCfg {
  static constexpr int RB = KC * 2;                         // bytes per voxel row of the halo
  static constexpr int HALO_ROWS = 180 * (TD + 2);          // 10 x 18 x (TD+2)
  static constexpr int HALO_TX = HALO_ROWS * RB;
  static constexpr int HALO_BYTES = (HALO_TX + 1023) / 1024 * 1024;
  static constexpr int NHALO = 2;
  static constexpr int TPB = BN <= 64 ? 3 : 1;              // taps per weight TMA box (the kd = 0,1,2 tiles of one (kh,kw))
  static constexpr int B_TAP = BN * KC * 2;                 // bytes of one tap's weight tile
  static constexpr int B_BOX = TPB * B_TAP;                 // bytes of one weight TMA box
  static constexpr int B_BOX_BYTES = (B_BOX + 1023) / 1024 * 1024;
  // (kh,kw) boxes per weight stage (template parameter KW, chosen per launch by launch_conv_halo).  KW = 3: a stage holds the
  // three kw boxes of one kh: 36 instead of 12 MMAs between two stage hand-backs (the wait / fence / descriptor set-up of a
  // hand-back cannot overlap the MMAs of the same issuing thread, ~150 cycles each -- tools/umma_rate.py: 76 cycles per MMA
  // at 12 per stage vs 64 back to back).  Needs 1024-byte multiples per box (TMA destination / UMMA descriptor alignment).
  static constexpr int KWS = KW;
  static_assert(KW == 1 || (KW == 3 && BN <= 32 && B_BOX % 1024 == 0), "three-box stages: BN <= 32, 1 KB aligned boxes");
  static constexpr int B_TX = KWS * B_BOX;
  static constexpr int B_BYTES = (B_TX + 1023) / 1024 * 1024;
  static constexpr int NB_MAX = 24;
  static constexpr int CBO = BN < 64 ? BN : 64;             // channels per output staging box
  static constexpr int NBO = BN / CBO;
  static constexpr int OUT_BOX = 128 * CBO * 2;
  static constexpr int OUT_TILE = 128 * BN * 2;             // one plane, one of {hi, lo}
  static constexpr int NACC = (2 * NI * TD * BN <= 512) ? 2 : 1;
  static constexpr int ACC_COLS = NACC * NI * TD * BN;
  static_assert(NI == 1 || NI == 2, "one or two issuing warps");
  static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static constexpr int AUX_BYTES = 1024 + 8 * BN * 2 * 4 + BN * 16;   // barriers | per-warp stats | GN coefficients
  static constexpr int BUDGET = 232448 - 1024;                   // dynamic smem limit minus alignment slack
  static constexpr uint32_t LAYOUT = KC == 64 ? UMMA_SW128 : KC == 32 ? UMMA_SW64 : UMMA_SW32;
  static constexpr uint32_t SBO_A = 10 * RB;                // one halo row (10 voxels) per 8-row group
  static constexpr uint32_t SBO_B = 8 * RB;
  static constexpr bool STK = BN <= 64;                     // kd taps stacked along N (one MMA feeds <= 3 output planes)
  static constexpr bool COLSPLIT = BN >= 64;                // epilogue groups split the columns (else the planes)
  static constexpr bool RUN = BN <= 64;                     // register-resident running statistics (<= 32 columns per thread)
  static_assert(ACC_COLS <= 512, "TMEM budget exceeded");
  static_assert((2 * NHALO + 2 * NB_MAX + 2 * NACC) * 8 + 8 <= 768, "barrier area overflow");   // side_full[6] sits at +768
};

// division by a run-time constant as multiply-high + shift (q = umulhi(x, mul) >> shr for x < 2^31; CUTLASS FastDivmod scheme):
// the five div/mod pairs of the tile decode cost ~700 cycles per tile in the epilogue warps (tools/halo_timeline.py)
struct FastDiv {
  unsigned mul, shr, div;
};
static inline FastDiv make_fastdiv(unsigned d) {
  FastDiv f;
  f.div = d;
  if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
  unsigned l = 31u - (unsigned)__builtin_clz(d);
  if (d & (d - 1)) ++l;                       // ceil(log2(d))
  const unsigned p = 31u + l;
  f.mul = (unsigned)((((unsigned long long)1 << p) + d - 1) / d);
  f.shr = p - 32u;
  return f;
}
__device__ __forceinline__ void fast_divmod(int x, const FastDiv& f, int& q, int& r) {
  q = f.div > 1 ? (int)(__umulhi((unsigned)x, f.mul) >> f.shr) : x;
  r = x - q * (int)f.div;
}

static constexpr int HALO_NO_RING = -1000;   // internal: the side-ring staging buffers do not fit this configuration

struct HaloArgs {
  int tiles_total;   // N * tiles_d * tiles_h * tiles_w * ntiles
  int ntiles;        // output-channel tiles
  int hsplit;        // the halo box is loaded as (TD+2) * hsplit TMA boxes of (KC, 10, 18/hsplit, 1)
  int nb;            // weight ring depth (stages of TPB taps)
  int nout;          // output staging buffers (2 or 4), each OUT_TILE * (split ? 2 : 1) bytes
  int split;
  FastDiv fd_nt, fd_w, fd_h, fd_d;   // ntiles, tiles_w, tiles_h, tiles_d
  int dense1;        // 1x1x1 sources are loaded as one dense (KC, 8, 16, TD) box instead of a halo neighbourhood
  int prefetch;      // L2-prefetch the later planes' side-input rows at tile start (measured: the prefetch instructions themselves
                     // stall the issuing warp ~450 cycles each; B200UNET_HALO_PREFETCH=1 re-enables them)
  int side_ring;     // lean epilogues: the side input (residual / GroupNorm input) of a plane is TMA-loaded INTO the plane's output
                     // staging buffer two planes ahead (maps.side) and transformed in place; needs 3 staging buffers per group
  long long* dbg;    // optional timeline buffer [3 roles][32 tiles][4] of clock64 stamps written by CTA 0 (tuning aid)
};
#define EPI_STAMP(idx) \
  do { if (hp.dbg && blockIdx.x == 0 && ti == 4 && warp == 2 && lane == 0) hp.dbg[384 + (idx)] = clock64(); } while (0)
#define HALO_STAMP(role, slot) \
  do { if (hp.dbg && blockIdx.x == 0 && ti < 32 && lane == 0) hp.dbg[((role) * 32 + ti) * 4 + (slot)] = clock64(); } while (0)

// EV (epilogue variant): 0 = every option at run time; 1 = mode 0 in single-pass bf16 without dropout scale / bias / zero
// boundary; 2 = mode 1 (GroupNorm-backward epilogue) in single-pass bf16.  The lean variants drop the dead branches from the
// unrolled epilogue body: ncu's source view attributed ~20-25 % of the epilogue warps' stall samples to instruction fetch
// (stall_no_inst) in the one-size-fits-all body (profiles/r02_halo_epilogue_stalls.txt).
template <int KC, int BN, int TD, int NI = 1, int KW = 1, int EV = 0>
__global__ void __launch_bounds__(NI == 2 ? 384 : 352, 1) k_conv_halo(const __grid_constant__ ConvMaps maps, const ConvArgs p,
                                                                      const HaloArgs hp) {
  using Cfg = HaloCfg<KC, BN, TD, NI, KW>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: an integer round trip loses the address space and every
  // shared-memory access below would compile to a generic LD.E / ST.E (ncu source view, round 2) instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int out_buf_bytes = Cfg::OUT_TILE * (hp.split ? 2 : 1);
  uint8_t* smem_halo = smem;
  uint8_t* smem_out = smem + Cfg::NHALO * Cfg::HALO_BYTES;
  uint8_t* smem_b = smem_out + hp.nout * out_buf_bytes;
  uint8_t* aux = smem_b + hp.nb * Cfg::B_BYTES;
  uint64_t* halo_full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* halo_empty = halo_full + Cfg::NHALO;
  uint64_t* b_full = halo_empty + Cfg::NHALO;
  uint64_t* b_empty = b_full + Cfg::NB_MAX;
  uint64_t* acc_full = b_empty + Cfg::NB_MAX;
  uint64_t* acc_empty = acc_full + Cfg::NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + Cfg::NACC);
  uint64_t* side_full = reinterpret_cast<uint64_t*>(aux + 768);   // [6]: one per staging buffer in side-ring mode
  float* s_stats = reinterpret_cast<float*>(aux + 1024);
  float4* s_coef = reinterpret_cast<float4*>(aux + 1024 + 8 * BN * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t NB = hp.nb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0][0]);
    tma_prefetch_desc(&maps.b[0][0]);
    tma_prefetch_desc(&maps.o[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::NHALO; ++s) { mbar_init(&halo_full[s], 1); mbar_init(&halo_empty[s], NI); }
      for (int s = 0; s < Cfg::NB_MAX; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
      for (int s = 0; s < Cfg::NACC; ++s) { mbar_init(&acc_full[s], NI); mbar_init(&acc_empty[s], 8); }
      for (int s = 0; s < 6; ++s) mbar_init(&side_full[s], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // everything above overlapped the tail of the previous kernel; nothing below may
  pdl_launch_dependents();

  // K groups of one tile: source 0 = 27 taps per chunk (9 stages of TPB taps when TPB = 3), source 1 (optional
  // 1x1x1) = its centre tap per chunk
  const int groups0 = p.kchunks[0], groups1 = p.ntaps[1] ? p.kchunks[1] : 0;
  constexpr int STAGES0 = 27 / (Cfg::TPB * Cfg::KWS);

  if (warp == 0 || warp == 10) {
    // ------------------------------------------------------------------ TMA producers (convergent, one lane issues)
    // warp 10 loads the halo boxes, warp 0 the weight stages: with a single producer the next halo could only be
    // requested after the last weight stage of the current chunk had been issued (<= nb stages before it is needed),
    // which left the MMA warp waiting ~1K cycles at every chunk / tile boundary (HALO_STAMP timeline).
    const bool halo_role = warp == 10;
    const uint32_t issue = elect_one() ? 1u : 0u;
    uint32_t hi = 0, bi = 0, ti = 0;
    for (int tile = blockIdx.x; tile < hp.tiles_total; tile += gridDim.x, ++ti) {
      int t = tile;
      int nt, wt, ht, dt, n;
      fast_divmod(t, hp.fd_nt, t, nt);
      fast_divmod(t, hp.fd_w, t, wt);
      fast_divmod(t, hp.fd_h, t, ht);
      fast_divmod(t, hp.fd_d, n, dt);
      const int w0 = wt * 8, h0 = ht * 16, d0 = dt * TD, n0 = nt * BN;
      if (halo_role) HALO_STAMP(0, 0);
      for (int g = 0; g < groups0 + groups1; ++g) {
        const int src = g < groups0 ? 0 : 1;
        const int kc = src == 0 ? g : g - groups0;
        const int nstage = src == 0 ? STAGES0 : 1;
        for (int pass = 0; pass < p.npass; ++pass) {
          if (halo_role) {
            const uint32_t s = hi % Cfg::NHALO, ph = (hi / Cfg::NHALO) & 1;
            mbar_wait(&halo_empty[s], ph ^ 1);
            HALO_STAMP(0, 1);
            if (src == 1 && hp.dense1) {
              // 1x1x1 source (fused `sample`, or a lone 1x1x1 convolution): no halo needed -- ONE dense box (KC, 8, 16, TD) in the
              // standard K-major layout (8-row groups 8 * RB apart, planes 128 rows apart) instead of the 10 x 18 x (TD+2)
              // neighbourhood: 2.1x less L2 -> shared-memory traffic for these sources
              mbar_expect_tx_if(issue, &halo_full[s], 128 * TD * Cfg::RB);
              tma_load_5d_if(issue, smem_halo + s * Cfg::HALO_BYTES, &maps.a[1][pass == 1], &halo_full[s], kc * KC, w0, h0, d0, n);
            } else {
              mbar_expect_tx_if(issue, &halo_full[s], Cfg::HALO_TX);
              const int hrows = 18 / hp.hsplit;
              for (int dp = 0; dp < TD + 2; ++dp)
                for (int hq = 0; hq < hp.hsplit; ++hq)
                  tma_load_5d_if(issue, smem_halo + s * Cfg::HALO_BYTES + ((dp * 18 + hq * hrows) * 10) * Cfg::RB,
                                 &maps.a[src][pass == 1], &halo_full[s], kc * KC, w0 - 1, h0 - 1 + hq * hrows, d0 - 1 + dp, n);
            }
            ++hi;
          } else {
            for (int st = 0; st < nstage; ++st) {
              const uint32_t s = bi % NB, ph = (bi / NB) & 1;
              mbar_wait(&b_empty[s], ph ^ 1);
              mbar_expect_tx_if(issue, &b_full[s], src == 0 ? Cfg::B_TX : Cfg::B_TAP);
              if (Cfg::STK && src == 0) {   // box = (kh,kw): the three kd taps through the 4-D (Cin, Cout, khkw, kd) view
#pragma unroll
                for (int q = 0; q < Cfg::KWS; ++q)
                  tma_load_4d_if(issue, smem_b + s * Cfg::B_BYTES + q * Cfg::B_BOX, &maps.b[src][pass == 2], &b_full[s], kc * KC, n0,
                                 st * Cfg::KWS + q, 0);
              } else
                tma_load_3d_if(issue, smem_b + s * Cfg::B_BYTES, &maps.b[src][pass == 2], &b_full[s], kc * KC, n0,
                               src == 0 ? st * Cfg::TPB : 0);
              ++bi;
            }
          }
        }
      }
    }
  } else if (warp == 1 || (NI == 2 && warp == 11)) {
    // ------------------------------------------------------------------ MMA issuer (one elected lane per stage)
    // tcgen05.mma issue is the critical path once the MMAs run near their operand-fetch bound of (128 + N) / 4 cycles:
    // a UTCHMMA holds its uniform-register operands until the tensor pipe dequeues it, so the wait / fence / descriptor
    // set-up of the next stage cannot run ahead (tools/umma_rate.py: one issuing warp reaches 95 cycles per N=128 MMA at
    // 4 MMAs per stage, 76 at 12).  Two issuing warps alternating stages reach the bound in the micro-benchmark, but
    // MMAs of DIFFERENT threads accumulating into the same TMEM columns are not ordered: tools/conv_determinism.py
    // showed lost updates (thousands of elements differing run to run), so with NI = 1 a single thread issues every MMA
    // of a tile.  NI = 2 (experimental): warp `iw` owns the stages of parity iw and accumulates them into its own
    // accumulator set; no column is ever written by two threads and the epilogue adds the two sets.
    // The stage loop is rolled with incremental tap offsets (a 27-stage unrolled body thrashed the instruction cache);
    // the weight-ring slot and phase are carried instead of recomputed with div/mod.
    const uint32_t iw = (NI == 2 && warp == 11) ? 1u : 0u;
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t idesc2 = make_idesc_bf16(128, BN * 2 <= 256 ? BN * 2 : BN, 0, 0);
    constexpr uint32_t idesc3 = make_idesc_bf16(128, BN * 3 <= 256 ? BN * 3 : BN, 0, 0);
    constexpr uint32_t hi_a = desc_hi(Cfg::SBO_A, Cfg::LAYOUT);
    constexpr uint32_t hi_b = desc_hi(Cfg::SBO_B, Cfg::LAYOUT);
    const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t halo0 = smem_u32(smem_halo), b0 = smem_u32(smem_b);
    uint32_t hs = 0, hph = 0, bs = 0, bph = 0, ti = 0;   // halo / weight ring slot and phase
    for (int tile = blockIdx.x; tile < hp.tiles_total; tile += gridDim.x, ++ti) {
      const uint32_t as = ti % Cfg::NACC;
      if (iw == 0) HALO_STAMP(1, 0);
      mbar_wait(&acc_empty[as], ((ti / Cfg::NACC) & 1) ^ 1);
      tc_fence_after();
      if (iw == 0) HALO_STAMP(1, 1);
      const uint32_t acc0 = tmem0 + (as * NI + iw) * TD * BN;
      uint32_t first = 1, sidx = 0;   // first: this warp has not issued into its accumulator set yet; sidx: stage index in the tile
      for (int g = 0; g < groups0 + groups1; ++g) {
        const int src = g < groups0 ? 0 : 1;
        for (int pass = 0; pass < p.npass; ++pass) {
          mbar_wait(&halo_full[hs], hph);
          if (iw == 0) HALO_STAMP(1, 2);
          const uint32_t halo_lo = desc_lo(halo0 + hs * Cfg::HALO_BYTES, 16);
          if (src == 0) {
            // rolled on purpose: the 27-stage unrolled body (~4K instructions) thrashed the instruction cache
            // (stall_no_inst was the top MMA-warp stall in the ncu source view); tap offsets advance incrementally
            uint32_t a_off = 0;          // descriptor offset ((kd*18 + kh)*10 + kw) * RB >> 4 of the stage's first tap
            int kw = 0, kh = 0;
#pragma unroll 1
            for (int st = 0; st < STAGES0; ++st) {
              if (NI == 1 || ((sidx + st) & 1u) == iw) {
                mbar_wait(&b_full[bs], bph);
                tc_fence_after();
                const uint32_t b_lo0 = desc_lo(b0 + bs * Cfg::B_BYTES, 16);
                const uint32_t a_base = halo_lo + a_off;
                if (elect_one()) {   // one elected lane issues the whole stage (ptxas keeps the block in uniform registers)
                  if constexpr (Cfg::STK) {
                    // kd-stacked issue: stage st = (kh,kw) holds the weight tiles of kd = 0,1,2 back to back (3*BN
                    // rows).  Halo plane hq feeds output planes hq-kd; accumulators sit in DESCENDING plane order in
                    // TMEM, so one MMA with N = nkd*BN columns starting at plane (hq-kdmin) covers them: N = 32 costs
                    // 40-46 cycles, N = 96 only 56, i.e. 6 MMAs replace 12 per (kh,kw,k16) at TD = 4.
                    static_assert(Cfg::TPB == 3, "stacked boxes hold 3 taps");
#pragma unroll
                    for (int q = 0; q < Cfg::KWS; ++q) {   // KWS = 3: the stage holds the kw = 0,1,2 boxes of one kh
                      const uint32_t a_q = a_base + ((q * Cfg::RB) >> 4);
                      const uint32_t b_q = b_lo0 + ((q * Cfg::B_BOX) >> 4);
#pragma unroll
                      for (int k = 0; k < KC / 16; ++k) {
                        if (q == 0 && k == 0 && first) {
                          // very first K step of the tile: unstacked, the kd = 0 MMA of every plane overwrites
#pragma unroll
                          for (int dpl = 0; dpl < TD; ++dpl) {
#pragma unroll
                            for (int kd = 0; kd < 3; ++kd)
                              umma_bf16(acc0 + (TD - 1 - dpl) * BN, desc_from(a_q + (((dpl + kd) * 180 * Cfg::RB) >> 4), hi_a),
                                        desc_from(b_q + ((kd * Cfg::B_TAP) >> 4), hi_b), idesc, kd > 0 ? 1u : 0u);
                          }
                        } else {
#pragma unroll
                          for (int hq = 0; hq < TD + 2; ++hq) {
                            const int kdmin = hq - (TD - 1) > 0 ? hq - (TD - 1) : 0;
                            const int kdmax = hq < 2 ? hq : 2;
                            const int nkd = kdmax - kdmin + 1;
                            const uint32_t idn = nkd == 1 ? idesc : nkd == 2 ? idesc2 : idesc3;
                            umma_bf16(acc0 + (TD - 1 - hq + kdmin) * BN, desc_from(a_q + ((hq * 180 * Cfg::RB + k * 32) >> 4), hi_a),
                                      desc_from(b_q + ((kdmin * Cfg::B_TAP + k * 32) >> 4), hi_b), idn, 1u);
                          }
                        }
                      }
                    }
                  } else {
                    static_assert(Cfg::STK || Cfg::TPB == 1, "unstacked stages hold one tap");
#pragma unroll
                    for (int dpl = 0; dpl < TD; ++dpl) {
#pragma unroll
                      for (int k = 0; k < KC / 16; ++k)
                        umma_bf16(acc0 + dpl * BN, desc_from(a_base + ((dpl * 180 * Cfg::RB + k * 32) >> 4), hi_a),
                                  desc_from(b_lo0 + ((k * 32) >> 4), hi_b), idesc, k == 0 ? (first ^ 1u) : 1u);
                    }
                  }
                  umma_commit(&b_empty[bs]);
                }
                __syncwarp();
                first = 0;
              }
              if (++bs == NB) { bs = 0; bph ^= 1; }
              if constexpr (Cfg::KWS == 3) {
                a_off += (10 * Cfg::RB) >> 4;   // next stage = next kh: one halo row of 10 voxels further
              } else {
                // next tap: kw fastest, then kh, (then kd for unstacked stages)
                a_off += Cfg::RB >> 4;
                if (++kw == 3) {
                  kw = 0;
                  a_off += (7 * Cfg::RB) >> 4;
                  if (++kh == 3) { kh = 0; a_off += (15 * 10 * Cfg::RB) >> 4; }
                }
              }
            }
            sidx += STAGES0;
          } else {
            // fused 1x1x1 source: one stage holding its only tap, read at the halo centre (kd = kh = kw = 1)
            if (NI == 1 || (sidx & 1u) == iw) {
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_lo0 = desc_lo(b0 + bs * Cfg::B_BYTES, 16);
            // dense box: plane dpl starts 128 rows further, standard K-major 8-row-group stride; halo layout: the centre voxel
            // (kd = kh = kw = 1) of the 10 x 18 x (TD+2) neighbourhood, 10-voxel row pitch
            const uint32_t a_lo = hp.dense1 ? halo_lo : halo_lo + ((((1 * 18 + 1) * 10 + 1) * Cfg::RB) >> 4);
            const uint32_t hi_a1 = hp.dense1 ? desc_hi(8 * Cfg::RB, Cfg::LAYOUT) : hi_a;
            const uint32_t pstride = (hp.dense1 ? 128 : 180) * Cfg::RB;
            if (elect_one()) {
#pragma unroll
            for (int dpl = 0; dpl < TD; ++dpl) {
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                umma_bf16(acc0 + (Cfg::STK ? TD - 1 - dpl : dpl) * BN,
                             desc_from(a_lo + ((dpl * pstride + k * 32) >> 4), hi_a1), desc_from(b_lo0 + ((k * 32) >> 4), hi_b),
                             idesc, k == 0 ? (first ^ 1u) : 1u);   // first: a lone 1x1x1 convolution starts the tile here
            }
            umma_commit(&b_empty[bs]);
            }
            __syncwarp();
            first = 0;
            }
            ++sidx;
            if (++bs == NB) { bs = 0; bph ^= 1; }
          }
          if (elect_one()) umma_commit(&halo_empty[hs]);
          __syncwarp();
          if (++hs == Cfg::NHALO) { hs = 0; hph ^= 1; }
        }
      }
      if (elect_one()) umma_commit(&acc_full[as]);
      __syncwarp();
      if (iw == 0) HALO_STAMP(1, 3);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // 8 warps in two groups of 4 (one warp per TMEM lane quadrant in each group), so two drains are in flight per SM
    // and every scheduler interleaves two epilogue warps (a single warp per scheduler ran latency-bound at ~600 cycles
    // per 16-column chunk: profiles/r01_halo_timeline.txt).
    //   BN <= 32 (plane split): group g drains the planes dpl = g, g+2, ... into its own staging buffers.
    //   BN >= 64 (column split): both groups drain every plane, group g the column half [g*BN/2, (g+1)*BN/2), into a
    //   shared staging buffer; a thread then owns <= 32 (BN=64) columns, so the running statistics fit in registers.
    // Side inputs (residual / norm input) are software-pipelined: the rows of the NEXT 32-column group (of this plane or
    // of the next one) are requested as soon as the registers of the current group have been consumed.
    constexpr bool RUN = Cfg::RUN;
    constexpr bool COLS = Cfg::COLSPLIT;
    constexpr int NJ = COLS ? BN / 32 : BN / 16;      // 16-column chunks a thread drains per plane
    constexpr int NACCUM = RUN ? NJ * 16 : 16;
    constexpr int SG = NJ < 2 ? NJ : 2;               // chunks per side-input group
    constexpr int NSG = NJ / SG;
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const int e = threadIdx.x - 64;           // 0..255
    const int grp = (warp - 2) >> 2;
    const int jb = COLS ? grp * NJ : 0;       // first chunk of this thread
    const int issuer = COLS ? 64 : 64 + grp * 128;   // the thread that issues the TMA stores (of its group)
    const int nog = COLS ? hp.nout : hp.nout >> 1;   // staging buffers this thread's group rotates through
    const int pstep = COLS ? 1 : 2;
    const int mode = EV == 1 ? 0 : EV == 2 ? 1 : p.mode;           // compile-time constants in the lean variants
    const bool split = EV != 0 ? false : hp.split != 0;
    const float* const scale = EV != 0 ? nullptr : p.scale;
    const float* const bias = EV != 0 ? nullptr : p.bias;
    const int zero_last = EV != 0 ? 0 : p.zero_last;
    const bool want_stats = (mode == 0) ? (p.stats != nullptr) : (p.bstats != nullptr);
    const bf16* side_hi = mode == 0 ? p.res_hi : p.x_hi;
    const bf16* side_lo = EV != 0 ? nullptr : (mode == 0 ? p.res_lo : p.x_lo);
    const int side_ld = mode == 0 ? p.ldr : p.ldx;
    double* stat_dst = (mode == 0) ? p.stats : p.bstats;
    const int stat_ld = (mode == 0) ? p.stats_ld : p.coef_ld;
    // Side-ring mode (lean variants with a side input; the host dispatches the generic variant when the three staging buffers
    // per group do not fit).  The side rows used to be read by every thread straight from global memory into registers, one
    // plane ahead at most: the ncu source view of the 32->64 GroupNorm-backward launch (profiles/r02_halo_mode1_source.txt) put
    // 15 % of ALL samples on the first instruction that waits for those loads and 11 % on the group barrier behind it.  Now
    // the group's issuing thread TMA-loads the side tile of the plane two planes ahead INTO that plane's output staging
    // buffer (same box and swizzle as the store), the threads transform their own 16-byte chunks in place, and the buffer
    // is TMA-stored: no registers, no per-thread global loads, ~1.5 plane times of latency cover.
    const bool ring = EV != 0 && side_hi != nullptr;
    const int ppg = COLS ? TD : (TD - grp + 1) / 2;      // planes of one tile this group drains
    const int sbar0 = COLS ? 0 : grp * 3;                // this group's side_full barriers (one per staging buffer, <= 3)
    auto issue_side = [&](uint32_t q) {                  // issuing thread: request plane q of the group's plane sequence
      const int tq = (int)(q / (uint32_t)ppg), kk = (int)(q % (uint32_t)ppg);
      const long long tl = (long long)blockIdx.x + (long long)tq * gridDim.x;
      if (tl >= hp.tiles_total) return;
      int t = (int)tl, nt, wt, ht, dt, n;
      fast_divmod(t, hp.fd_nt, t, nt);
      fast_divmod(t, hp.fd_w, t, wt);
      fast_divmod(t, hp.fd_h, t, ht);
      fast_divmod(t, hp.fd_d, n, dt);
      const int d = dt * TD + (COLS ? 0 : grp) + kk * pstep;     // may lie beyond Do: the box then reads as zeros
      uint8_t* dst = smem_out + ((COLS ? 0 : grp * nog) + (q % (uint32_t)nog)) * out_buf_bytes;
      uint64_t* bar = &side_full[sbar0 + (q % (uint32_t)nog)];
      int nbx = 0;
#pragma unroll
      for (int cb = 0; cb < Cfg::NBO; ++cb) nbx += (nt * BN + cb * Cfg::CBO < p.Cout) ? 1 : 0;
      mbar_expect_tx(bar, nbx * Cfg::OUT_BOX);
#pragma unroll
      for (int cb = 0; cb < Cfg::NBO; ++cb)
        if (nt * BN + cb * Cfg::CBO < p.Cout)
          tma_load_5d(dst + cb * Cfg::OUT_BOX, &maps.side, bar, nt * BN + cb * Cfg::CBO, wt * 8, ht * 16, d, n);
    };
    float rs[NACCUM], rq[NACCUM];
#pragma unroll
    for (int i = 0; i < NACCUM; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
    for (int i = e; i < 8 * BN * 2; i += 256) s_stats[i] = 0.f;
    asm volatile("bar.sync 3, 256;" ::: "memory");
    uint32_t ti = 0, oi = 0;   // tile counter, output-plane counter (staging ring)
    int cur_n = -1, cur_n0 = -1;
    auto group_sync = [&]() {
      if (COLS) asm volatile("bar.sync 3, 256;" ::: "memory");
      else if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
    };

    // s_stats holds one private [BN][2] slot per epilogue warp (shared-memory float atomics are CAS spin loops and
    // collapse under 8-warp contention); the flush sums the slots -> global fp64 atomics, then re-zeroes them
    float2* s_mine = reinterpret_cast<float2*>(s_stats) + (warp - 2) * BN;
    auto flush_smem = [&](int fn, int fn0) {
      asm volatile("bar.sync 3, 256;" ::: "memory");
      for (int c = e; c < BN * 2; c += 256) {
        float v = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) { v += s_stats[wq * BN * 2 + c]; s_stats[wq * BN * 2 + c] = 0.f; }
        if (fn0 + (c >> 1) < p.Cout) atomicAdd(&stat_dst[((long long)fn * stat_ld + fn0) * 2 + c], (double)v);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
    };
    // register accumulators of chunk j -> transposing butterfly -> this warp's s_stats slot
    auto reduce_chunk = [&](int j, const float* sv, const float* sq) {
      float v16[16], q16[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { v16[i] = sv[i]; q16[i] = sq[i]; }
      const float s1 = warp_colsum16(v16, lane), s2 = warp_colsum16(q16, lane);
      if ((lane & 1) == 0) {
        const int col = j * 16 + ((lane >> 1) & 15);
        float2 acc = s_mine[col];
        if constexpr (EV == 2) {   // raw sum(dz * x) -> sum(dz * xhat) with this run's (mean, rstd) of the channel
          const float* f = reinterpret_cast<const float*>(s_coef) + (col >> 1) * 8 + (col & 1);
          acc.x += s1; acc.y += f[6] * (s2 - f[4] * s1);
        } else {
          acc.x += s1; acc.y += s2;
        }
        s_mine[col] = acc;
      }
    };
    // side-input rows of chunk group sg (SG chunks of 16 channels) of the voxel row at vx
    uint4 sh[2 * SG], sl[2 * SG];
    auto load_side = [&](long long vx, int n0, int sg, bool ok) {
      if constexpr (EV == 0) {                       // lean variants: side ring (above)
        if (!side_hi || !ok) return;
#pragma unroll
        for (int c = 0; c < 2 * SG; ++c) {
          const int cc = n0 + (jb + sg * SG) * 16 + c * 8;
          if (cc < p.Cout) {
            sh[c] = *reinterpret_cast<const uint4*>(side_hi + vx * side_ld + cc);
            if (side_lo) sl[c] = *reinterpret_cast<const uint4*>(side_lo + vx * side_ld + cc);
          }
        }
      }
    };

    // look-ahead = buffers per group - 1: two planes with three buffers; one plane with two (configurations whose three-box
    // weight stages leave no room for a third buffer)
    const uint32_t la = (uint32_t)nog - 1u;
    if (ring && threadIdx.x == issuer && ppg > 0)
      for (uint32_t q = 0; q < la; ++q) issue_side(q);
    for (int tile = blockIdx.x; tile < hp.tiles_total; tile += gridDim.x, ++ti) {
      EPI_STAMP(40);
      int t = tile;
      int nt, wt, ht, dt, n;
      fast_divmod(t, hp.fd_nt, t, nt);
      fast_divmod(t, hp.fd_w, t, wt);
      fast_divmod(t, hp.fd_h, t, ht);
      fast_divmod(t, hp.fd_d, n, dt);
      const int w0 = wt * 8, h0 = ht * 16, d0 = dt * TD, n0 = nt * BN;
      if (n != cur_n || n0 != cur_n0) {
        if (RUN && want_stats && cur_n >= 0) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) reduce_chunk(jb + j, rs + (RUN ? j * 16 : 0), rq + (RUN ? j * 16 : 0));
#pragma unroll
          for (int i = 0; i < NACCUM; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
          flush_smem(cur_n, cur_n0);
        }
        if (mode == 1) {
          asm volatile("bar.sync 3, 256;" ::: "memory");   // nobody still reads the old coefficients
          for (int c = e; c < BN; c += 256) {
            const float4 k = (n0 + c < p.Cout) ? p.coef[(long long)n * p.coef_ld + n0 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (EV == 2) {
              // pair-interleaved for the packed f32x2 arithmetic: channel pair q = c / 2 holds [a0 a1 b0 b1 | mean0 mean1 rstd0 rstd1]
              float* f = reinterpret_cast<float*>(s_coef) + (c >> 1) * 8 + (c & 1);
              f[0] = k.x; f[2] = k.y; f[4] = k.z; f[6] = k.w;
            } else {
              s_coef[c] = k;
            }
          }
          asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        cur_n = n; cur_n0 = n0;
      }
      const int w = w0 + (row & 7), h = h0 + (row >> 3);
      const bool valid_wh = (w < p.Wo) && (h < p.Ho);
      const long long vox0 = (((long long)n * p.Do + d0) * p.Ho + h) * p.Wo + w;   // plane dpl: + dpl * Ho * Wo
      const long long plane = (long long)p.Ho * p.Wo;
      const int dpl0 = COLS ? 0 : grp;
      EPI_STAMP(41);
      load_side(vox0 + dpl0 * plane, n0, 0, valid_wh && (d0 + dpl0 < p.Do) && dpl0 < TD);   // lands while the MMAs still run
      EPI_STAMP(42);
      if (hp.prefetch)
        for (int dpl = dpl0 + pstep; dpl < TD; dpl += pstep)     // the later planes' rows -> L2
          conv_epilogue_prefetch(p, n0 + jb * 16, NJ * 16, vox0 + dpl * plane, valid_wh && (d0 + dpl < p.Do));
      EPI_STAMP(43);
      const uint32_t as = ti % Cfg::NACC;
      if (warp == 2) HALO_STAMP(2, 0);
      mbar_wait(&acc_full[as], (ti / Cfg::NACC) & 1);
      tc_fence_after();
      if (warp == 2) HALO_STAMP(2, 1);
      EPI_STAMP(44);
#pragma unroll 1
      for (int dpl = dpl0; dpl < TD; dpl += pstep, ++oi) {
        const int d = d0 + dpl;
        const bool valid = valid_wh && (d < p.Do);
        const long long vox = vox0 + dpl * plane;
        const bool valid_next = valid_wh && (dpl + pstep < TD) && (d + pstep < p.Do);
        uint8_t* stage = smem_out + ((COLS ? 0 : grp * nog) + (oi % nog)) * out_buf_bytes;
        EPI_STAMP(dpl * 8 + 0);
        if (nog == 1) {   // single staging buffer: its previous store must have finished reading it
          if (threadIdx.x == issuer) tma_store_wait_read0();
          group_sync();
        }
        if (ring) mbar_wait(&side_full[sbar0 + oi % (uint32_t)nog], (oi / (uint32_t)nog) & 1u);   // this plane's side tile sits in `stage`
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) {
          const int j = jb + jj;
          const int c0 = n0 + j * 16;
          float* as_ = rs + (RUN ? jj * 16 : 0);
          float* aq_ = rq + (RUN ? jj * 16 : 0);
          if (c0 < p.Cout) {
            uint32_t r[16];
            const uint32_t tcol = (as * NI * TD + (Cfg::STK ? TD - 1 - dpl : dpl)) * BN + j * 16;
            tmem_ld16(tmem_base + tcol + (static_cast<uint32_t>(lane_base) << 16), r);
            if constexpr (NI == 2) {   // the second issuer's partial sums live TD planes further
              uint32_t r2[16];
              tmem_ld16(tmem_base + tcol + TD * BN + (static_cast<uint32_t>(lane_base) << 16), r2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
            }
            tmem_ld_wait();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const int cc = c0 + hf * 8;
              // the 8 channels (16 bytes) of this row in the staging tile: box (cc - n0) / CBO, chunk within the box row
              const int cb = (j * 16 + hf * 8) / Cfg::CBO, cchunk = ((j * 16 + hf * 8) % Cfg::CBO) / 8;
              uint8_t* dst = stage + cb * Cfg::OUT_BOX + stage_off<Cfg::CBO>(row, cchunk);
              float vv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) vv[i] = 0.f;
              if (cc < p.Cout && valid) {
                float sv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) vv[i] = __uint_as_float(r[hf * 8 + i]);
                if (side_hi) {
                  const uint4 a = EV != 0 ? *reinterpret_cast<const uint4*>(dst) : sh[(jj % SG) * 2 + hf];
                  sv[0] = bf16_lo_to_f(a.x); sv[1] = bf16_hi_to_f(a.x); sv[2] = bf16_lo_to_f(a.y); sv[3] = bf16_hi_to_f(a.y);
                  sv[4] = bf16_lo_to_f(a.z); sv[5] = bf16_hi_to_f(a.z); sv[6] = bf16_lo_to_f(a.w); sv[7] = bf16_hi_to_f(a.w);
                  if (side_lo) {
                    const uint4 b = sl[(jj % SG) * 2 + hf];
                    sv[0] += bf16_lo_to_f(b.x); sv[1] += bf16_hi_to_f(b.x); sv[2] += bf16_lo_to_f(b.y); sv[3] += bf16_hi_to_f(b.y);
                    sv[4] += bf16_lo_to_f(b.z); sv[5] += bf16_hi_to_f(b.z); sv[6] += bf16_lo_to_f(b.w); sv[7] += bf16_hi_to_f(b.w);
                  }
                }
                if (mode == 0) {
                  if (side_hi) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vv[i] += sv[i];
                  }
                  if (scale) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vv[i] *= __ldg(scale + (long long)n * p.Cout + cc + i);
                  }
                  if (bias) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vv[i] += __ldg(bias + cc + i);
                  }
                  if (zero_last && (w == p.Wo - 1 || h == p.Ho - 1 || d == p.Do - 1)) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) vv[i] = 0.f;
                  }
#pragma unroll
                  for (int i = 0; i < 8; ++i) { as_[hf * 8 + i] += vv[i]; aq_[hf * 8 + i] = fmaf(vv[i], vv[i], aq_[hf * 8 + i]); }
                } else if constexpr (EV == 2) {
                  // packed f32x2 arithmetic on channel pairs; the second statistic is accumulated as the RAW sum(dz * x) and
                  // centred / scaled once per run in reduce_chunk: sum(dz * xhat) = rstd * (sum(dz * x) - mean * sum(dz))
#pragma unroll
                  for (int i = 0; i < 8; i += 2) {
                    const float4 ab = s_coef[(j * 16 + hf * 8 + i)];          // [a_i a_i+1 b_i b_i+1] (pair-interleaved, see above)
                    const float2 x2 = make_float2(sv[i], sv[i + 1]);
                    const float2 z2 = __ffma2_rn(make_float2(ab.x, ab.y), x2, make_float2(ab.z, ab.w));
                    const float2 m2 = make_float2(z2.x > 0.f ? 1.f : p.slope, z2.y > 0.f ? 1.f : p.slope);
                    const float2 dz2 = __fmul2_rn(make_float2(vv[i], vv[i + 1]), m2);
                    vv[i] = dz2.x; vv[i + 1] = dz2.y;
                    const float2 s2 = __fadd2_rn(make_float2(as_[hf * 8 + i], as_[hf * 8 + i + 1]), dz2);
                    as_[hf * 8 + i] = s2.x; as_[hf * 8 + i + 1] = s2.y;
                    const float2 q2 = __ffma2_rn(dz2, x2, make_float2(aq_[hf * 8 + i], aq_[hf * 8 + i + 1]));
                    aq_[hf * 8 + i] = q2.x; aq_[hf * 8 + i + 1] = q2.y;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float4 k = s_coef[j * 16 + hf * 8 + i];
                    const float z = fmaf(k.x, sv[i], k.y);
                    const float dz = z > 0.f ? vv[i] : vv[i] * p.slope;
                    vv[i] = dz;
                    as_[hf * 8 + i] += dz;
                    aq_[hf * 8 + i] = fmaf(dz, (sv[i] - k.z) * k.w, aq_[hf * 8 + i]);
                  }
                }
              }
              uint4 o;
              o.x = pack_bf16x2(vv[0], vv[1]); o.y = pack_bf16x2(vv[2], vv[3]);
              o.z = pack_bf16x2(vv[4], vv[5]); o.w = pack_bf16x2(vv[6], vv[7]);
              *reinterpret_cast<uint4*>(dst) = o;
              if (split) {
                uint4 l;
                l.x = pack_bf16x2(vv[0] - bf16_lo_to_f(o.x), vv[1] - bf16_hi_to_f(o.x));
                l.y = pack_bf16x2(vv[2] - bf16_lo_to_f(o.y), vv[3] - bf16_hi_to_f(o.y));
                l.z = pack_bf16x2(vv[4] - bf16_lo_to_f(o.z), vv[5] - bf16_hi_to_f(o.z));
                l.w = pack_bf16x2(vv[6] - bf16_lo_to_f(o.w), vv[7] - bf16_hi_to_f(o.w));
                *reinterpret_cast<uint4*>(dst + Cfg::OUT_TILE) = l;
              }
            }
            if constexpr (!RUN) {
              if (want_stats) reduce_chunk(j, rs, rq);
#pragma unroll
              for (int i = 0; i < 16; ++i) { rs[i] = 0.f; rq[i] = 0.f; }
            }
          }
          if (ring && jj == 0 && threadIdx.x == issuer) {
            // buffer (oi + la) % nog was stored from at the end of the previous plane: once that store has read it, the side tile
            // of the plane `la` ahead can land there
            tma_store_wait_read0();
            issue_side(oi + la);
          }
          // the registers of this side-input group are consumed: request the next group (this plane's, else the next plane's)
          if ((jj % SG) == SG - 1) {
            if (jj / SG + 1 < NSG) load_side(vox, n0, jj / SG + 1, valid);
            else load_side(vox + pstep * plane, n0, 0, valid_next);
          }
        }
        // the plane is staged: make it visible to the async proxy, then one thread TMA-stores it
        EPI_STAMP(dpl * 8 + 2);
        fence_proxy_async();
        EPI_STAMP(dpl * 8 + 3);
        if (!ring && nog > 1 && threadIdx.x == issuer) tma_store_wait_read0();   // the other buffer is free again after the barrier
        EPI_STAMP(dpl * 8 + 4);
        group_sync();
        EPI_STAMP(dpl * 8 + 5);
        if (threadIdx.x == issuer && d < p.Do) {
#pragma unroll
          for (int cb = 0; cb < Cfg::NBO; ++cb) {
            if (n0 + cb * Cfg::CBO < p.Cout) {
              tma_store_5d(&maps.o[0], stage + cb * Cfg::OUT_BOX, n0 + cb * Cfg::CBO, w0, h0, d, n);
              if (split) tma_store_5d(&maps.o[1], stage + Cfg::OUT_TILE + cb * Cfg::OUT_BOX, n0 + cb * Cfg::CBO, w0, h0, d, n);
            }
          }
          tma_store_commit();
        }
        EPI_STAMP(dpl * 8 + 6);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);      // accumulator set drained: MMA may overwrite it
      EPI_STAMP(45);
      if (warp == 2) HALO_STAMP(2, 2);
      if (!RUN && want_stats) flush_smem(n, n0);
      if (warp == 2) HALO_STAMP(2, 3);
    }
    if (RUN && want_stats && cur_n >= 0) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) reduce_chunk(jb + j, rs + (RUN ? j * 16 : 0), rq + (RUN ? j * 16 : 0));
      flush_smem(cur_n, cur_n0);
    }
    if (threadIdx.x == issuer) tma_store_wait_all();   // all output tiles have left shared memory and are written
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int KC, int BN, int TD, int NI = 1, int KW = 1, int EV = 0>
static int launch_halo_cfg(const ConvMaps& maps, const ConvArgs& a, HaloArgs h, int grid, cudaStream_t st) {
  using Cfg = HaloCfg<KC, BN, TD, NI, KW>;
  // shared-memory carve-up: 2 halo buffers | nout output staging buffers | weight ring | aux
  const int out_buf = Cfg::OUT_TILE * (h.split ? 2 : 1);
  const int rem = Cfg::BUDGET - Cfg::AUX_BYTES - Cfg::NHALO * Cfg::HALO_BYTES;
  // staging buffers: column-split groups share a ring of 2; plane-split groups own 2 each if they fit, else 1 each
  h.nout = Cfg::COLSPLIT ? 2 : (rem - 4 * out_buf >= 4 * Cfg::B_BOX_BYTES && rem - 4 * out_buf >= 2 * Cfg::B_BYTES) ? 4 : 2;
  if (const char* e = getenv("B200UNET_HALO_NOUT")) {   // tuning override
    const int v = atoi(e);
    if (Cfg::COLSPLIT) { if (v == 1 || v == 2) h.nout = v; }   // shared ring of 1 or 2
    else if (v == 2 || v == 4) h.nout = v;                      // 1 or 2 per plane-split group
  }
  if (h.side_ring) {   // side-ring mode: three staging buffers per epilogue group (two-plane look-ahead) and still >= 3 weight
    h.nout = Cfg::COLSPLIT ? 3 : 6;     // stages; else two per group (one-plane look-ahead) for the plane-split tiles
    if ((rem - h.nout * out_buf) / Cfg::B_BYTES < 3) {
      h.nout = 4;
      if (Cfg::COLSPLIT || (rem - h.nout * out_buf) / Cfg::B_BYTES < 3) return HALO_NO_RING;   // the caller falls back
    }
  }
  int nb = (rem - h.nout * out_buf) / Cfg::B_BYTES;
  if (nb > Cfg::NB_MAX) nb = Cfg::NB_MAX;
  B200_REQUIRE(nb >= 2, E_UNSUPPORTED, "conv_halo: configuration KC=%d BN=%d TD=%d does not fit shared memory", KC, BN, TD);
  h.nb = nb;
  const int smem_bytes = Cfg::NHALO * Cfg::HALO_BYTES + h.nout * out_buf + nb * Cfg::B_BYTES + Cfg::AUX_BYTES + 1024;
  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(k_conv_halo<KC, BN, TD, NI, KW, EV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set[dev] = true;
  }
  launch_pdl(k_conv_halo<KC, BN, TD, NI, KW, EV>, dim3(grid), dim3(NI == 2 ? 384 : 352), smem_bytes, st, maps, a, h);
  B200_CHECK_CUDA(cudaGetLastError());
  return OK;
}


// every (KC, BN, TD, KW) combination of one epilogue variant; returns E_UNSUPPORTED when there is none
template <int EV>
static int launch_halo_table(int KC, int BN, int TD, int kws, const ConvMaps& maps, const ConvArgs& a, const HaloArgs& h, int grid,
                             cudaStream_t st) {
  if (kws == 3) {
#define B200_HALO_CASE3(kc, bn, td) \
  if (KC == kc && BN == bn && TD == td) return launch_halo_cfg<kc, bn, td, 1, 3, EV>(maps, a, h, grid, st);
    B200_HALO_CASE3(16, 32, 4) B200_HALO_CASE3(16, 32, 2) B200_HALO_CASE3(16, 32, 1)
    B200_HALO_CASE3(32, 16, 4) B200_HALO_CASE3(32, 16, 2) B200_HALO_CASE3(32, 16, 1)
    B200_HALO_CASE3(32, 32, 4) B200_HALO_CASE3(32, 32, 2) B200_HALO_CASE3(32, 32, 1)
#undef B200_HALO_CASE3
  }
#define B200_HALO_CASE(kc, bn, td) \
  if (KC == kc && BN == bn && TD == td) return launch_halo_cfg<kc, bn, td, 1, 1, EV>(maps, a, h, grid, st);
  B200_HALO_CASE(16, 16, 4) B200_HALO_CASE(16, 16, 2) B200_HALO_CASE(16, 16, 1)
  B200_HALO_CASE(16, 32, 4) B200_HALO_CASE(16, 32, 2) B200_HALO_CASE(16, 32, 1)
  B200_HALO_CASE(16, 64, 4) B200_HALO_CASE(16, 64, 2) B200_HALO_CASE(16, 64, 1)
  B200_HALO_CASE(16, 128, 4) B200_HALO_CASE(16, 128, 2) B200_HALO_CASE(16, 128, 1)
  B200_HALO_CASE(32, 16, 4) B200_HALO_CASE(32, 16, 2) B200_HALO_CASE(32, 16, 1)
  B200_HALO_CASE(32, 32, 4) B200_HALO_CASE(32, 32, 2) B200_HALO_CASE(32, 32, 1)
  B200_HALO_CASE(32, 64, 4) B200_HALO_CASE(32, 64, 2) B200_HALO_CASE(32, 64, 1)
  B200_HALO_CASE(32, 128, 4) B200_HALO_CASE(32, 128, 2) B200_HALO_CASE(32, 128, 1)
#undef B200_HALO_CASE
  set_error("conv_halo: no kernel for KC=%d BN=%d TD=%d", KC, BN, TD);
  return E_UNSUPPORTED;
}

}  // namespace b200
