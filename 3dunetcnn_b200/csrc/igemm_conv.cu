// 3-D convolution as an implicit GEMM on the sm_100a tensor cores.
//
//   Y[v][co] = sum_{tap, ci} A[v*stride + tap - pad][ci] * Wp[tap][co][ci]          (fp32 accumulate in TMEM)
//
// Replaces (forward, and data-gradient with flipped/transposed packed weights):
//   nn.Conv3d k3 s1/s2 p1, bias-free    /root/reference/unet3d/models/pytorch/classification/resnet.py:12-17
//   nn.Conv3d k1                        /root/reference/unet3d/models/pytorch/classification/resnet.py:20-22
// GEMM view: M = 128 output voxels (one tw x th x td spatial box of one sample), N = BN output channels,
// K = taps x Cin walked in chunks of KC channels.  Per K step the producer thread issues two TMA loads:
//   A: 5-D box (KC, tw, th, td, 1) of the NDHWC activation at the tap-shifted coordinate; the zero padding of the
//      convolution is TMA out-of-bounds fill, stride-2 convolutions use the tensor map's element strides;
//   B: 3-D box (KC, BN, 1) of the packed weights [tap][co][ci].
// Both land K-major with the hardware swizzle that matches KC (128B/64B/32B) and feed tcgen05.mma
// (cta_group::1, kind::f16, M=128, N=BN, K=16) issued by one thread; a STAGES-deep mbarrier ring decouples
// TMA from MMA, tcgen05.commit releases ring slots and finally signals the epilogue warps.
// Epilogue (4 warps, one TMEM lane quadrant each): tcgen05.ld -> registers ->
//   mode 0: (+ residual) (* per-(n,c) dropout scale) -> bf16 hi[/lo] store, per-channel sum / sum-of-squares
//           for the next GroupNorm (warp butterfly -> smem -> one double atomic per channel per CTA);
//   mode 1: GroupNorm/ReLU backward: dz = dact * 1[A x + B > 0], per-channel (sum dz, sum dz*xhat).
// Split-precision ("parity") mode runs three passes per K step: Ah*Wh, Al*Wh, Ah*Wl.
#include <cstdlib>
#include "conv_common.cuh"

namespace b200 {

// DEEP: the pipeline ring takes ~196 KB instead of 96 KB.  The 96 KB ring lets two CTAs share an SM (192 KB of loads in flight per
// SM); a launch with no more CTAs than SMs has one CTA per SM whatever it allocates, and with three 32 KB stages in flight it is
// bound by the L2 round trip (the 256-channel layers at 16^3: ~34 B/clk per SM, profiles/r02_layer_times.csv) -- those launches
// take the deep ring.
template <int BN, int KC, int DEEP = 0>
struct ConvCfg {
  static constexpr int A_BYTES = 128 * KC * 2;
  static constexpr int B_BOX_BYTES = BN * KC * 2;
  static constexpr int B_BYTES = B_BOX_BYTES < 1024 ? 1024 : B_BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = ((DEEP ? 196 : 96) * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
  static constexpr int AUX_BYTES = 1024 + 4 * BN * 2 * 4 + BN * 16;  // barriers+slot | per-warp stats | coef
  static constexpr int OUT_STAGING = 2 * 128 * BN * 2;    // hi + lo output tiles, aliased onto the pipeline stages
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES > OUT_STAGING ? STAGES * STAGE_BYTES : OUT_STAGING;
  static constexpr int SMEM_BYTES = PIPE_BYTES + AUX_BYTES + 1024;  // +1024 alignment slack
  static constexpr uint32_t LAYOUT = KC == 64 ? UMMA_SW128 : KC == 32 ? UMMA_SW64 : UMMA_SW32;
  static constexpr uint32_t SBO = 8 * KC * 2;
};

template <int BN, int KC, int DEEP>
__global__ void __launch_bounds__(192) k_igemm_conv(const __grid_constant__ ConvMaps maps, const ConvArgs p,
                                                    const __grid_constant__ ConvClassMaps cmaps) {
  using Cfg = ConvCfg<BN, KC, DEEP>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: an integer round trip loses the address space and every
  // shared-memory access below would compile to a generic LD.E / ST.E (ncu source view, round 2) instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* aux = smem + Cfg::PIPE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  float* s_stats = reinterpret_cast<float*>(aux + 1024);             // [4 warps][BN][2]
  float4* s_coef = reinterpret_cast<float4*>(aux + 1024 + 4 * BN * 8);   // [BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  int t = blockIdx.x;
  const int wt = t % p.tiles_w; t /= p.tiles_w;
  const int ht = t % p.tiles_h; t /= p.tiles_h;
  const int dt = t % p.tiles_d;
  const int n = t / p.tiles_d;
  const int w0 = wt * p.tw, h0 = ht * p.th, d0 = dt * p.td;
  const int n0 = blockIdx.y * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0][0]);
    tma_prefetch_desc(&maps.b[0][0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(tfull_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, p.cls_mode ? (8 * BN < 32 ? 32 : 8 * BN) : Cfg::TMEM_COLS);   // class mode: one accumulator per parity class
    tmem_relinquish();
  }
  pdl_wait();   // before the first global read (the coefficient table below); barrier init / TMEM allocation above overlap the previous kernel
  if (warp >= 2) {
    const int e = threadIdx.x - 64;
    for (int i = e; i < 4 * BN * 2; i += 128) s_stats[i] = 0.f;
    if (p.mode == 1) {
      for (int c = e; c < BN; c += 128)
        s_coef[c] = (n0 + c < p.Cout) ? p.coef[(long long)n * p.coef_ld + n0 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  // class mode: ONE CTA computes all eight parity classes of its voxel tile (27 tap products walked class by class into
  // eight TMEM accumulators, then eight tile stores).  One CTA per class spent most of its life in set-up: 32768 CTAs of
  // 1-8 K steps each took 0.47 ms for the 32-channel level-0 gradient.
  int total_iters = (p.ntaps[0] * p.kchunks[0] + p.ntaps[1] * p.kchunks[1]) * p.npass;
  if (p.cls_mode) {
    total_iters = 0;
    for (int c = 0; c < 8; ++c) total_iters += (int)p.cls_n[c] * p.kchunks[0] * p.npass;
  }

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (convergent, one lane issues)
    {
      const uint32_t issue = elect_one() ? 1u : 0u;
      int it = 0;
      for (int src = 0; src < 2; ++src) {
        const int nt = p.ntaps[src];
        if (nt == 0) continue;
        const int ks = p.ksz[src], pad = p.pad[src], sd = p.stride[src];
        for (int cls = 0; cls < (p.cls_mode ? 8 : 1); ++cls) {
        const int ntap_loop = p.cls_mode ? (int)p.cls_n[cls] : nt;
        for (int ti = 0; ti < ntap_loop; ++ti) {
          int tap = ti, cw, ch, cd;
          if (p.cls_mode) {   // class tap list: source voxel j + delta, packed-weight tap index
            const int e = p.cls_tap[cls][ti];
            tap = e & 31;
            cw = w0 + ((e >> 5) & 1); ch = h0 + ((e >> 6) & 1); cd = d0 + ((e >> 7) & 1);
          } else {
            const int kd = tap / (ks * ks), kh = (tap / ks) % ks, kw = tap % ks;
            cw = w0 * sd + kw - pad; ch = h0 * sd + kh - pad; cd = d0 * sd + kd - pad;
          }
          for (int kc = 0; kc < p.kchunks[src]; ++kc) {
            for (int pass = 0; pass < p.npass; ++pass) {
              const int s = it % Cfg::STAGES;
              const uint32_t ph = (it / Cfg::STAGES) & 1;
              mbar_wait(&empty_bar[s], ph ^ 1);
              mbar_expect_tx_if(issue, &full_bar[s], Cfg::A_BYTES + Cfg::B_BOX_BYTES);
              uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
              tma_load_5d_if(issue, sa, &maps.a[src][pass == 1], &full_bar[s], kc * KC, cw, ch, cd, n);
              tma_load_3d_if(issue, sa + Cfg::A_BYTES, &maps.b[src][pass == 2], &full_bar[s], kc * KC, n0, tap);
              ++it;
            }
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (convergent, one lane issues)
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t hi_d = desc_hi(Cfg::SBO, Cfg::LAYOUT);
      const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem0 = smem_u32(smem);
      // class mode: iterations [cls_end[c-1], cls_end[c]) accumulate into accumulator c (TMEM columns c * BN ...)
      int cur_cls = 0, cls_end = p.cls_mode ? (int)p.cls_n[0] * p.kchunks[0] * p.npass : total_iters, cls_first = 0;
      for (int it = 0; it < total_iters; ++it) {
        while (it >= cls_end) { ++cur_cls; cls_first = it; cls_end += (int)p.cls_n[cur_cls] * p.kchunks[0] * p.npass; }
        const int s = it % Cfg::STAGES;
        const uint32_t ph = (it / Cfg::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_lo = desc_lo(smem0 + s * Cfg::STAGE_BYTES, 16);
        const uint32_t b_lo = desc_lo(smem0 + s * Cfg::STAGE_BYTES + Cfg::A_BYTES, 16);
        const uint32_t acc = tmem0 + cur_cls * BN;
        if (elect_one()) {   // one elected lane issues the stage (descriptors stay in uniform registers)
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            umma_bf16(acc, desc_from(a_lo + 2 * k, hi_d), desc_from(b_lo + 2 * k, hi_d), idesc, (it > cls_first || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (it == total_iters - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const int wl = row % p.tw, hl = (row / p.tw) % p.th, dl = row / (p.tw * p.th);
    const int w = w0 + wl, h = h0 + hl, d = d0 + dl;
    const bool valid = (w < p.Wo) && (h < p.Ho) && (d < p.Do);
    const bool want_stats = (p.mode == 0) ? (p.stats != nullptr) : (p.bstats != nullptr);
    const bool edge = p.zero_last && (w == p.Wo - 1 || h == p.Ho - 1 || d == p.Do - 1);
    const bool split = p.out_lo != nullptr;
    constexpr int CBO = BN < 64 ? BN : 64;
    // side inputs (residual) are indexed in the OUTPUT tensor: in class mode that is voxel 2j + p of a 2x grid
    auto vox_of = [&](int cls) -> long long {
      return p.cls_mode
          ? (((long long)n * (2 * p.Do) + 2 * d + ((cls >> 2) & 1)) * (2 * p.Ho) + 2 * h + ((cls >> 1) & 1)) * (2 * p.Wo) + 2 * w + (cls & 1)
          : (((long long)n * p.Do + d) * p.Ho + h) * p.Wo + w;
    };
    for (int cls = 0; cls < (p.cls_mode ? 8 : 1); ++cls) conv_epilogue_prefetch(p, n0, BN, vox_of(cls), valid);   // -> L2 while the MMAs run
    for (int cls = 0; cls < (p.cls_mode ? 8 : 1); ++cls) {
      const long long vox = vox_of(cls);
      // cls_pair: classes (pd, ph, 0) and (pd, ph, 1) interleave along W in ONE staging tile of 256 rows (row = output voxel
      // (dl, hl, 2 wl + pw) of a box 2 tw wide) that is stored through the dense class-pair map after the second drain
      const bool pair = p.cls_pair != 0, pair_first = pair && (cls & 1) == 0, pair_second = pair && (cls & 1) == 1;
      if (cls == 0) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // s_stats / s_coef initialised
        mbar_wait(tfull_bar, 0);
        tc_fence_after();
      } else if (!pair_second) {
        if (threadIdx.x == 64) tma_store_wait_read0();    // the previous stores have read the staging tile
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const int srow = pair ? (dl * p.th + hl) * (2 * p.tw) + 2 * wl + (cls & 1) : row;
      constexpr int SROWS_BOX = 128 * CBO * 2;
      // all MMAs have completed (tfull) => every pipeline stage has been consumed: the stage memory is free and is reused
      // as the output staging tile [BN/CBO boxes][128 (pair: 256) rows][CBO] (+ lo tile), TMA-stored below
      conv_epilogue_tile<BN>(p, tmem_base + cls * BN, lane_base, lane, n, n0, vox, valid, s_stats, s_coef, want_stats, edge, smem, srow, split);
      if (pair_first) continue;     // the other W parity fills the odd rows of the same tile
      fence_proxy_async();
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) {
        if (pair) {
          tma_store_5d(&cmaps.oc[cls & 6][0], smem, n0, 2 * w0, h0, d0, n);
        } else {
          const CUtensorMap* mo_hi = p.cls_mode ? &cmaps.oc[cls][0] : &maps.o[0];
          const CUtensorMap* mo_lo = p.cls_mode ? &cmaps.oc[cls][1] : &maps.o[1];
#pragma unroll
          for (int cb = 0; cb < BN / CBO; ++cb) {
            if (n0 + cb * CBO < p.Cout) {
              tma_store_5d(mo_hi, smem + cb * SROWS_BOX, n0 + cb * CBO, w0, h0, d0, n);
              if (split) tma_store_5d(mo_lo, smem + 128 * BN * 2 + cb * SROWS_BOX, n0 + cb * CBO, w0, h0, d0, n);
            }
          }
        }
        tma_store_commit();
      }
    }
    if (want_stats) {
      const int e = threadIdx.x - 64;
      double* dst = (p.mode == 0) ? p.stats : p.bstats;
      const int ld = (p.mode == 0) ? p.stats_ld : p.coef_ld;
      for (int c = e; c < BN * 2; c += 128) {
        const float v = s_stats[c] + s_stats[BN * 2 + c] + s_stats[2 * BN * 2 + c] + s_stats[3 * BN * 2 + c];
        if (n0 + (c >> 1) < p.Cout) atomicAdd(&dst[((long long)n * ld + n0) * 2 + c], (double)v);
      }
    }
    if (threadIdx.x == 64) tma_store_wait_all();   // the staging tile must outlive the bulk store
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.cls_mode ? (8 * BN < 32 ? 32 : 8 * BN) : Cfg::TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------- host side
static void pick_tile(int Wo, int Ho, int Do, int& tw, int& th, int& td) {
  tw = Wo >= 8 ? 8 : Wo >= 4 ? 4 : Wo >= 2 ? 2 : 1;
  int rem = 128 / tw;
  th = Ho >= 4 ? 4 : Ho >= 2 ? 2 : 1;
  if (th > rem) th = rem;
  td = rem / th;
}

template <int BN, int KC, int DEEP = 0>
static int launch_cfg(const ConvMaps& maps, const ConvArgs& args, dim3 grid, cudaStream_t st, const ConvClassMaps& cmaps) {
  using Cfg = ConvCfg<BN, KC, DEEP>;
  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(k_igemm_conv<BN, KC, DEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  launch_pdl(k_igemm_conv<BN, KC, DEEP>, grid, dim3(192), Cfg::SMEM_BYTES, st, maps, args, cmaps);
  B200_CHECK_CUDA(cudaGetLastError());
  return OK;
}

static const int kIgemmDeepDefault = 1;   // measured: -12 % on the 256-channel 16^3 launches (profiles/r02_igemm_deep_ab.txt); B200UNET_IGEMM_DEEP=0 restores the 96 KB ring

static int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return 148;
  if (!cached[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

int launch_igemm_conv(const ConvOp& op, cudaStream_t st) {
  static const bool no_halo = getenv("B200UNET_NO_HALO") != nullptr;
  if (!no_halo && !op.cls_mode && (op.nsrc == 1 || op.nsrc == 2) && conv_halo_eligible(op)) return launch_conv_halo(op, sm_count(), st);
  return launch_igemm_conv_streaming(op, st);
}

int launch_igemm_conv_streaming(const ConvOp& op, cudaStream_t st) {
  B200_REQUIRE(op.nsrc == 1 || op.nsrc == 2, E_INVALID, "igemm_conv: nsrc=%d", op.nsrc);
  const Act& out = op.out;
  B200_REQUIRE(out.C % 8 == 0 && out.ld % 8 == 0, E_UNSUPPORTED, "igemm_conv: Cout=%d (pitch %d) must be a multiple of 8",
               out.C, out.ld);
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  // class mode: the GEMM rows are the voxels of ONE parity class of the output = the source grid
  const int gD = op.cls_mode ? out.D / 2 : out.D, gH = op.cls_mode ? out.H / 2 : out.H, gW = op.cls_mode ? out.W / 2 : out.W;
  a.N = out.N; a.Do = gD; a.Ho = gH; a.Wo = gW; a.Cout = out.C;
  pick_tile(gW, gH, gD, a.tw, a.th, a.td);
  a.tiles_w = ceil_div(gW, a.tw); a.tiles_h = ceil_div(gH, a.th); a.tiles_d = ceil_div(gD, a.td);
  int cin_max = 0;
  bool split = false;
  for (int s = 0; s < op.nsrc; ++s) {
    const ConvSrc& c = op.src[s];
    B200_REQUIRE(c.ksz == 1 || c.ksz == 3 || (c.ksz == 2 && c.nopad && (c.stride == 2 || op.cls_mode == 2)), E_UNSUPPORTED,
                 "igemm_conv: kernel_size=%d unsupported", c.ksz);
    B200_REQUIRE(c.stride == 1 || c.stride == 2, E_UNSUPPORTED, "igemm_conv: stride=%d unsupported", c.stride);
    B200_REQUIRE(c.x.C % 8 == 0 && c.x.ld % 8 == 0, E_UNSUPPORTED, "igemm_conv: Cin=%d must be a multiple of 8", c.x.C);
    B200_REQUIRE(c.x.N == out.N, E_INVALID, "igemm_conv: batch mismatch");
    const int pad = c.nopad ? 0 : c.ksz / 2;
    if (op.cls_mode)
      B200_REQUIRE(op.nsrc == 1 && c.ksz == (op.cls_mode == 2 ? 2 : 3) && 2 * c.x.D == out.D && 2 * c.x.H == out.H && 2 * c.x.W == out.W &&
                       op.mode == 0 && !op.bias && !op.zero_last,
                   E_INVALID, "igemm_conv: class mode needs one source at half the output extent and a plain epilogue");
    else
    B200_REQUIRE((c.x.D + 2 * pad - c.ksz) / c.stride + 1 == out.D && (c.x.H + 2 * pad - c.ksz) / c.stride + 1 == out.H &&
                     (c.x.W + 2 * pad - c.ksz) / c.stride + 1 == out.W,
                 E_INVALID, "igemm_conv: source %d dims %dx%dx%d (k%d s%d) do not produce output %dx%dx%d", s, c.x.D,
                 c.x.H, c.x.W, c.ksz, c.stride, out.D, out.H, out.W);
    if (c.x.C > cin_max) cin_max = c.x.C;
    if (c.x.lo || c.w_lo) split = true;
  }
  if (split) {
    for (int s = 0; s < op.nsrc; ++s)
      B200_REQUIRE(op.src[s].x.lo && op.src[s].w_lo, E_INVALID, "igemm_conv: split mode needs lo parts on every source");
  }
  const int KC = cin_max > 32 ? 64 : cin_max > 16 ? 32 : 16;
  int BN = out.C > 64 ? 128 : out.C > 32 ? 64 : out.C > 16 ? 32 : 16;
  if (op.cls_mode && BN > 64) BN = 64;   // eight accumulators of BN columns share the 512 TMEM columns
  const Swz swz = swz_for_bytes(KC * 2);
  for (int s = 0; s < op.nsrc; ++s) {
    const ConvSrc& c = op.src[s];
    a.ntaps[s] = c.ksz * c.ksz * c.ksz; a.ksz[s] = c.ksz; a.stride[s] = c.stride; a.pad[s] = c.nopad ? 0 : c.ksz / 2;
    a.kchunks[s] = ceil_div(c.x.C, KC);
    const int estride = op.cls_mode ? 1 : c.stride;
    B200_TRY(make_act_map(&maps.a[s][0], c.x.hi, c.x.N, c.x.D, c.x.H, c.x.W, c.x.C, c.x.ld, KC, a.tw, a.th, a.td,
                          estride, swz, c.x.vD, c.x.vH, c.x.vW));
    B200_TRY(make_w_map(&maps.b[s][0], c.w_hi, a.ntaps[s], op.Cop, c.Cip, KC, BN, swz));
    if (split) {
      B200_TRY(make_act_map(&maps.a[s][1], c.x.lo, c.x.N, c.x.D, c.x.H, c.x.W, c.x.C, c.x.ld, KC, a.tw, a.th, a.td,
                            estride, swz, c.x.vD, c.x.vH, c.x.vW));
      B200_TRY(make_w_map(&maps.b[s][1], c.w_lo, a.ntaps[s], op.Cop, c.Cip, KC, BN, swz));
    }
  }
  a.npass = split ? 3 : 1;
  a.mode = op.mode;
  a.out_hi = out.hi; a.out_lo = out.lo; a.ldo = out.ld;
  if (split) B200_REQUIRE(out.lo != nullptr, E_INVALID, "igemm_conv: split mode needs a lo output");
  ConvClassMaps cmaps;
  memset(&cmaps, 0, sizeof(cmaps));
  if (op.cls_mode) {
    // data gradient of y[o] = sum_k x[2o + k - 1] w[k]: dx[2j] = dy[j] w[1]; dx[2j+1] = dy[j] w[2] + dy[j+1] w[0].  With the
    // flipped pack Wd[k'] = w[2 - k'] (what emit_dgrad binds): even outputs use k' = 1 (delta 0), odd outputs k' = 0
    // (delta 0) and k' = 2 (delta +1); dy[j+1] beyond the grid reads as zero (TMA out-of-bounds fill).
    // cls_mode 2 (ConvTranspose3d, kernel = stride = 2): out[2j + p] = x[j] w[p], one tap per class
    a.cls_mode = op.cls_mode;
    const int cbo = BN < 64 ? BN : 64;
    for (int cls = 0; cls < 8; ++cls) {
      const int pd = (cls >> 2) & 1, ph = (cls >> 1) & 1, pw = cls & 1;
      int n = 0;
      if (op.cls_mode == 2) a.cls_tap[cls][n++] = (unsigned char)cls;   // tap index kd*4 + kh*2 + kw = the class itself
      else
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const bool ok = (pd ? kd != 1 : kd == 1) && (ph ? kh != 1 : kh == 1) && (pw ? kw != 1 : kw == 1);
            if (!ok) continue;
            a.cls_tap[cls][n++] = (unsigned char)((kd * 9 + kh * 3 + kw) | ((kw == 2) << 5) | ((kh == 2) << 6) | ((kd == 2) << 7));
          }
      a.cls_n[cls] = (unsigned char)n;
      B200_TRY(make_act_map_class(&cmaps.oc[cls][0], out.hi, out.N, out.D, out.H, out.W, out.C, out.ld, pd, ph, pw, cbo, a.tw, a.th,
                                  a.td, swz_for_bytes(cbo * 2)));
      if (split)
        B200_TRY(make_act_map_class(&cmaps.oc[cls][1], out.lo, out.N, out.D, out.H, out.W, out.C, out.ld, pd, ph, pw, cbo, a.tw,
                                    a.th, a.td, swz_for_bytes(cbo * 2)));
    }
    // single-pass bf16: the two W-parity classes of a (pd, ph) pair share one dense store (see tmap.h); the pair maps replace the
    // even classes' descriptors
    static const bool no_pair = getenv("B200UNET_CLASS_PAIR") && atoi(getenv("B200UNET_CLASS_PAIR")) == 0;   // A/B switch
    if (!split && !no_pair) {
      for (int cls = 0; cls < 8; cls += 2)
        B200_TRY(make_act_map_classpair(&cmaps.oc[cls][0], out.hi, out.N, out.D, out.H, out.W, out.C, out.ld, (cls >> 2) & 1, (cls >> 1) & 1, cbo,
                                        2 * a.tw, a.th, a.td, swz_for_bytes(cbo * 2)));
      a.cls_pair = 1;
    }
    maps.o[0] = cmaps.oc[0][0];   // keeps the (unused) plain output descriptor valid
  } else {
    const int cbo = BN < 64 ? BN : 64;
    B200_TRY(make_act_map(&maps.o[0], out.hi, out.N, out.D, out.H, out.W, out.C, out.ld, cbo, a.tw, a.th, a.td, 1,
                          swz_for_bytes(cbo * 2)));
    if (split)
      B200_TRY(make_act_map(&maps.o[1], out.lo, out.N, out.D, out.H, out.W, out.C, out.ld, cbo, a.tw, a.th, a.td, 1,
                            swz_for_bytes(cbo * 2)));
  }
  if (op.res) {
    B200_REQUIRE(op.res->C == out.C, E_INVALID, "igemm_conv: residual channel mismatch");
    a.res_hi = op.res->hi; a.res_lo = op.res->lo; a.ldr = op.res->ld;
  }
  a.scale = op.scale;
  a.bias = op.bias; a.zero_last = op.zero_last;
  a.stats = op.stats; a.stats_ld = op.stats_ld;
  if (op.mode == 1) {
    B200_REQUIRE(op.gn_x && op.coef, E_INVALID, "igemm_conv: mode 1 needs gn_x and coef");
    B200_REQUIRE(op.gn_x->C == out.C, E_INVALID, "igemm_conv: gn_x channel mismatch");
    a.x_hi = op.gn_x->hi; a.x_lo = op.gn_x->lo; a.ldx = op.gn_x->ld;
    a.coef = reinterpret_cast<const float4*>(op.coef); a.coef_ld = op.coef_ld;
    a.slope = op.slope; a.bstats = op.bstats;
  }
  dim3 grid((unsigned)((long long)a.N * a.tiles_d * a.tiles_h * a.tiles_w), (unsigned)ceil_div(out.C, BN), 1u);
  // at most one CTA per SM: nothing is lost by taking the whole shared memory for a deeper ring (see ConvCfg); only the 64-channel
  // K chunks have stages large enough for the 96 KB budget to cap the ring below six
  static const int deep_env = getenv("B200UNET_IGEMM_DEEP") ? atoi(getenv("B200UNET_IGEMM_DEEP")) : kIgemmDeepDefault;
  const bool deep = deep_env != 0 && !op.cls_mode && KC == 64 && (long long)grid.x * grid.y <= sm_count();
  if (deep && BN == 128) return launch_cfg<128, 64, 1>(maps, a, grid, st, cmaps);
  if (deep && BN == 64) return launch_cfg<64, 64, 1>(maps, a, grid, st, cmaps);
#define B200_CONV_CASE(bn, kc) \
  if (BN == bn && KC == kc) return launch_cfg<bn, kc>(maps, a, grid, st, cmaps);
  B200_CONV_CASE(16, 16) B200_CONV_CASE(16, 32) B200_CONV_CASE(16, 64)
  B200_CONV_CASE(32, 16) B200_CONV_CASE(32, 32) B200_CONV_CASE(32, 64)
  B200_CONV_CASE(64, 16) B200_CONV_CASE(64, 32) B200_CONV_CASE(64, 64)
  B200_CONV_CASE(128, 16) B200_CONV_CASE(128, 32) B200_CONV_CASE(128, 64)
#undef B200_CONV_CASE
  set_error("igemm_conv: no kernel for BN=%d KC=%d", BN, KC);
  return E_UNSUPPORTED;
}

}  // namespace b200
