// Halo-resident weight gradient (3x3x3, stride 1) for sm_100a: successor of wgrad.cu's per-tap streaming kernel
// for layers whose output plane is at least 8 x 16.
//
//   dW[tap][ci][co] += sum_v A[v + tap - 1][ci] * dY[v][co]
//
// A CTA owns up to 512/BN accumulators (128 x BN fp32 in TMEM), each one (ci-chunk kc, kd, kh) "row group" whose 128
// rows are 128/KC swizzle atoms along M = the taps kw = 0,1,2(,+unused) of KC input channels: the atoms are the SAME
// shared-memory halo box read at start addresses one voxel row apart (descriptor LBO = one row), so one TMA box
// (KC, 10, 18, TD+2) of the activation per voxel tile serves all 27 taps instead of 27 shifted boxes.
// K = voxels: one MMA contracts 16 voxels = 8(w) x 2(h): the two 8-row K groups are SBO = one halo row (A) / one
// dense tile row (dY) apart.  Split-K over voxel tiles across CTAs; fp32 vector atomics into dW at the end.
// Both operands are MN-major; shifted / re-strided descriptors rely on the absolute-address swizzle (probe.cu).
// The kd taps are stacked along N: halo plane hq (input depth d0-1+hq) meets output plane hq-kd for kd = 0,1,2, i.e. the
// dY planes hq-2, hq-1, hq, which lie one plane apart in shared memory -- the N swizzle atoms of one MMA (descriptor LBO
// = one dY plane).  An accumulator is therefore (kc, kh) x [kd = 2,1,0 blocks of BN columns], and one MMA of N = 3*BN
// (56 cycles at BN = 32) replaces three of N = BN (46 cycles each: the operand-fetch floor of (128 + N) / 4 cycles).
// The accumulators are zeroed once with tcgen05.st so that every MMA accumulates (an N-stacked MMA cannot overwrite only
// some of its column blocks).
#include <cstdlib>
#include "kernels.h"
#include "ptx.cuh"
#include "tmap.h"

namespace b200 {

struct WgHaloMaps {
  CUtensorMap a[2];
  CUtensorMap dy[2];
};

struct WgHaloArgs {
  int N, D, H, W;
  int Ci, Co, Cip, Cop;
  int tiles_w, tiles_h, tiles_d, tiles_total;
  int nkc;        // ci chunks
  int gpk;        // accumulator groups per ci chunk
  int qt;         // kh accumulators per CTA (<= 3), each 3*BN columns
  int splits;
  int npass;
  int hsplit;     // halo box loaded as (TD+2)*hsplit TMA boxes, dY tile as TD boxes (more requests in flight)
  float* dw;
  float* part;             // deterministic mode: per-split partial sums (plain stores) instead of atomics
  long long part_stride;
};

template <int KC, int BN, int TD>
struct WgHaloCfg {
  static constexpr int RB = KC * 2;
  static constexpr int HALO_TX = 180 * (TD + 2) * RB;
  static constexpr int HALO_BYTES = (HALO_TX + 1023) / 1024 * 1024;
  static constexpr int CBN = BN < 64 ? BN : 64;
  static constexpr int BPN = BN / CBN;
  static constexpr int RBN = CBN * 2;
  static constexpr int DY_BOX = 128 * TD * RBN;              // one box (CBN channels)
  static constexpr int DY_TX = DY_BOX * BPN;
  static constexpr int DY_BYTES = (DY_TX + 1023) / 1024 * 1024;
  static constexpr int NH = 2, ND = 2;
  static constexpr int SMEM_BYTES = NH * HALO_BYTES + ND * DY_BYTES + 1024 + 1024 + 1024;  // + tail slack + aux + align
  static constexpr uint32_t LAYOUT_A = KC == 64 ? UMMA_SW128 : KC == 32 ? UMMA_SW64 : UMMA_SW32;
  static constexpr uint32_t LAYOUT_B = CBN == 64 ? UMMA_SW128 : CBN == 32 ? UMMA_SW64 : UMMA_SW32;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

template <int KC, int BN, int TD>
__global__ void __launch_bounds__(192, 1) k_wgrad_halo(const __grid_constant__ WgHaloMaps maps, const WgHaloArgs p) {
  using Cfg = WgHaloCfg<KC, BN, TD>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: an integer round trip loses the address space and every
  // shared-memory access below would compile to a generic LD.E / ST.E (ncu source view, round 2) instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_h = smem;
  uint8_t* smem_d = smem + Cfg::NH * Cfg::HALO_BYTES + 1024;   // 1 KB slack: the unused 4th tap atom reads 3 rows past a box
  uint8_t* aux = smem_d + Cfg::ND * Cfg::DY_BYTES;
  uint64_t* h_full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* h_empty = h_full + Cfg::NH;
  uint64_t* d_full = h_empty + Cfg::NH;
  uint64_t* d_empty = d_full + Cfg::ND;
  uint64_t* tfull = d_empty + Cfg::ND;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // role: blockIdx.y = (kc, group) ; blockIdx.z = co tile ; blockIdx.x = split
  const int kc = blockIdx.y / p.gpk;
  const int grp = blockIdx.y % p.gpk;
  const int q0 = grp * p.qt;                       // first kh of this CTA
  const int nq = min(p.qt, 3 - q0);
  const int co0 = blockIdx.z * BN;
  const int t0 = (int)((long long)p.tiles_total * blockIdx.x / p.splits);
  const int t1 = (int)((long long)p.tiles_total * (blockIdx.x + 1) / p.splits);
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(nq * 3 * BN)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.dy[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::NH; ++s) { mbar_init(&h_full[s], 1); mbar_init(&h_empty[s], 1); }
      for (int s = 0; s < Cfg::ND; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], 1); }
      mbar_init(tfull, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // set-up above overlaps the tail of the previous kernel; no global access before this line
  pdl_launch_dependents();
  if (warp >= 2) {   // zero this warp's lane quadrant of every accumulator column
    for (int c = 0; c < nq * 3 * BN; c += 16) tmem_zero16(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + c);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    {
      const uint32_t issue = elect_one() ? 1u : 0u;
      uint32_t it = 0;
      for (int tile = t0; tile < t1; ++tile) {
        int t = tile;
        const int wt = t % p.tiles_w; t /= p.tiles_w;
        const int ht = t % p.tiles_h; t /= p.tiles_h;
        const int dt = t % p.tiles_d;
        const int n = t / p.tiles_d;
        const int w0 = wt * 8, h0 = ht * 16, d0 = dt * TD;
        for (int pass = 0; pass < p.npass; ++pass, ++it) {
          const uint32_t s = it % 2, ph = (it / 2) & 1;
          mbar_wait(&h_empty[s], ph ^ 1);
          mbar_expect_tx_if(issue, &h_full[s], Cfg::HALO_TX);
          const int hrows = 18 / p.hsplit;
          for (int dp = 0; dp < TD + 2; ++dp)
            for (int hq = 0; hq < p.hsplit; ++hq)
              tma_load_5d_if(issue, smem_h + s * Cfg::HALO_BYTES + ((dp * 18 + hq * hrows) * 10) * Cfg::RB, &maps.a[pass == 1],
                             &h_full[s], kc * KC, w0 - 1, h0 - 1 + hq * hrows, d0 - 1 + dp, n);
          mbar_wait(&d_empty[s], ph ^ 1);
          mbar_expect_tx_if(issue, &d_full[s], Cfg::DY_TX);
#pragma unroll
          for (int bx = 0; bx < Cfg::BPN; ++bx)
            for (int dp = 0; dp < TD; ++dp)
              tma_load_5d_if(issue, smem_d + s * Cfg::DY_BYTES + bx * Cfg::DY_BOX + dp * 128 * Cfg::RBN, &maps.dy[pass == 2],
                             &d_full[s], co0 + bx * Cfg::CBN, w0, h0, d0 + dp, n);
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc1 = make_idesc_bf16(128, BN, 1, 1);
      constexpr uint32_t idesc2 = make_idesc_bf16(128, 2 * BN, 1, 1);
      constexpr uint32_t idesc3 = make_idesc_bf16(128, 3 * BN <= 256 ? 3 * BN : BN, 1, 1);
      constexpr uint32_t hi_a = desc_hi(10 * Cfg::RB, Cfg::LAYOUT_A), hi_b = desc_hi(8 * Cfg::RBN, Cfg::LAYOUT_B);
      const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t h0s = smem_u32(smem_h), d0s = smem_u32(smem_d);
      uint32_t it = 0;
      for (int tile = t0; tile < t1; ++tile) {
        for (int pass = 0; pass < p.npass; ++pass, ++it) {
          const uint32_t s = it % 2, ph = (it / 2) & 1;
          mbar_wait(&h_full[s], ph);
          mbar_wait(&d_full[s], ph);
          tc_fence_after();
          const uint32_t h_lo = desc_lo(h0s + s * Cfg::HALO_BYTES, Cfg::RB);        // LBO = one voxel row: next kw tap
          const uint32_t d_lo = desc_lo(d0s + s * Cfg::DY_BYTES, 128 * Cfg::RBN);   // LBO = one dY plane: next kd block
          if (elect_one()) {   // one elected lane issues the whole tile pass (descriptors stay in uniform registers)
            for (int qi = 0; qi < nq; ++qi) {
              const int kh = q0 + qi;
              const uint32_t a_q = h_lo + ((kh * 10 * Cfg::RB) >> 4);
              const uint32_t acc = tmem0 + qi * 3 * BN;
#pragma unroll
              for (int hq = 0; hq < TD + 2; ++hq) {
                // halo plane hq pairs with output planes hq - kd, kd in [kdmin, kdmax]; column block (2 - kd), dY plane hq - kd
                const int kdmin = hq - (TD - 1) > 0 ? hq - (TD - 1) : 0;
                const int kdmax = hq < 2 ? hq : 2;
                const int nkd = kdmax - kdmin + 1;
                const uint32_t idn = nkd == 1 ? idesc1 : nkd == 2 ? idesc2 : idesc3;
#pragma unroll
                for (int hp = 0; hp < 8; ++hp)
                  umma_bf16(acc + (2 - kdmax) * BN, desc_from(a_q + ((((hq * 18 + 2 * hp) * 10) * Cfg::RB) >> 4), hi_a),
                            desc_from(d_lo + (((((hq - kdmax) * 16 + 2 * hp) * 8) * Cfg::RBN) >> 4), hi_b), idn, 1u);
              }
            }
            umma_commit(&h_empty[s]);
            umma_commit(&d_empty[s]);
          }
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit(tfull);
      __syncwarp();
    }
  } else {
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int kw = row / KC;
    const int ci = kc * KC + row % KC;
    const bool in_range = (kw < 3) && (ci < p.Ci);
    const bool has_work = t1 > t0;
    float* const base = p.part ? p.part + (long long)blockIdx.x * p.part_stride : p.dw;
    for (int qi = 0; qi < nq; ++qi) {
      for (int blk = 0; blk < 3; ++blk) {
        const int tap = ((2 - blk) * 3 + (q0 + qi)) * 3 + kw;   // column block blk holds kd = 2 - blk
#pragma unroll 1
        for (int j = 0; j < BN / 16; ++j) {
          uint32_t r[16];
          tmem_ld16(tmem_base + (static_cast<uint32_t>(lane_base) << 16) + (qi * 3 + blk) * BN + j * 16, r);
          tmem_ld_wait();
          const int c = co0 + j * 16;
          if (in_range && (has_work || p.part)) {
            float* dst = base + ((long long)tap * p.Cip + ci) * p.Cop + c;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              if (c + i + 3 < p.Cop) {
                float4 v = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                                       __uint_as_float(r[i + 3]));
                if (p.part) *reinterpret_cast<float4*>(dst + i) = has_work ? v : make_float4(0.f, 0.f, 0.f, 0.f);
                else atomicAdd(reinterpret_cast<float4*>(dst + i), v);
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

template <int KC, int BN, int TD>
static int launch_wgh(const WgHaloMaps& maps, const WgHaloArgs& a, dim3 grid, cudaStream_t st) {
  using Cfg = WgHaloCfg<KC, BN, TD>;
  static bool attr_set[64] = {false};
  int dev = 0;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(k_wgrad_halo<KC, BN, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES));
    attr_set[dev] = true;
  }
  launch_pdl(k_wgrad_halo<KC, BN, TD>, grid, dim3(192), Cfg::SMEM_BYTES, st, maps, a);
  B200_CHECK_CUDA(cudaGetLastError());
  return OK;
}

bool wgrad_halo_eligible(const WgradOp& op) {
  static const int max_c = getenv("B200UNET_WGHALO_MAXC") ? atoi(getenv("B200UNET_WGHALO_MAXC")) : 64;
  if (op.a.C > max_c) return false;
  return op.ksz == 3 && op.stride == 1 && op.dy.W >= 8 && op.dy.H >= 16;
}

int wgrad_halo_splits(const WgradOp& op, int num_sms) {
  const Act& A = op.a;
  const Act& Y = op.dy;
  const int KC = A.C > 16 ? 32 : 16;
  const int BN = Y.C > 32 ? 64 : Y.C > 16 ? 32 : 16;
  const int TD = (BN <= 32 && Y.D >= 4) ? 4 : 2;
  const int tiles_total = Y.N * ceil_div(Y.D, TD) * ceil_div(Y.H, 16) * ceil_div(Y.W, 8);
  int qt = 512 / (3 * BN);
  if (qt > 3) qt = 3;
  const int gpk = ceil_div(3, qt);
  const int roles = ceil_div(A.C, KC) * gpk * ceil_div(Y.C, BN);
  int splits = num_sms / roles;
  if (splits < 1) splits = 1;
  if (splits > tiles_total) splits = tiles_total;
  return splits;
}

int launch_wgrad_halo(const WgradOp& op, int num_sms, cudaStream_t st) {
  const Act& A = op.a;
  const Act& Y = op.dy;
  B200_REQUIRE(wgrad_halo_eligible(op), E_UNSUPPORTED, "wgrad_halo: shape not eligible");
  B200_REQUIRE(A.C % 8 == 0 && Y.C % 8 == 0 && A.ld % 8 == 0 && Y.ld % 8 == 0, E_UNSUPPORTED,
               "wgrad_halo: channels must be multiples of 8");
  B200_REQUIRE(op.Cop % 4 == 0 && op.Cop >= Y.C && op.Cip >= A.C, E_INVALID, "wgrad_halo: bad accumulator pitch");
  B200_REQUIRE(A.N == Y.N && A.D == Y.D && A.H == Y.H && A.W == Y.W, E_INVALID, "wgrad_halo: shape mismatch");
  const bool split = A.lo != nullptr || Y.lo != nullptr;
  if (split) B200_REQUIRE(A.lo && Y.lo, E_INVALID, "wgrad_halo: split mode needs lo parts on both operands");
  WgHaloArgs a;
  memset(&a, 0, sizeof(a));
  WgHaloMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int KC = A.C > 16 ? 32 : 16;
  int BN = Y.C > 32 ? 64 : Y.C > 16 ? 32 : 16;
  const int TD = (BN <= 32 && Y.D >= 4) ? 4 : 2;   // deeper tiles stack more kd taps per MMA (BN = 64: shared memory allows 2)
  const int CBN = BN < 64 ? BN : 64;
  a.N = Y.N; a.D = Y.D; a.H = Y.H; a.W = Y.W;
  a.Ci = A.C; a.Co = Y.C; a.Cip = op.Cip; a.Cop = op.Cop;
  a.tiles_w = ceil_div(Y.W, 8); a.tiles_h = ceil_div(Y.H, 16); a.tiles_d = ceil_div(Y.D, TD);
  a.tiles_total = a.N * a.tiles_d * a.tiles_h * a.tiles_w;
  a.nkc = ceil_div(A.C, KC);
  a.qt = 512 / (3 * BN);   // kh accumulators (3*BN columns each) per CTA
  if (a.qt > 3) a.qt = 3;
  a.gpk = ceil_div(3, a.qt);
  a.qt = ceil_div(3, a.gpk);
  const int cotiles = ceil_div(Y.C, BN);
  const int splits = wgrad_halo_splits(op, num_sms);
  a.splits = splits;
  a.npass = split ? 3 : 1;
  a.dw = op.dw;
  if (op.part) {
    a.part = op.part;
    a.part_stride = (long long)27 * op.Cip * op.Cop;
    B200_REQUIRE((size_t)splits * a.part_stride * sizeof(float) <= op.part_bytes, E_INVALID,
                 "wgrad_halo: deterministic partial buffer too small (%d splits)", splits);
    if (op.part_splits) *op.part_splits = splits;
  }
  int hsplit = KC == 32 ? 2 : 1;
  if (const char* e = getenv("B200UNET_HALO_HSPLIT")) {
    const int v = atoi(e);
    if ((v == 1 || v == 2 || v == 3 || v == 6) && ((18 / v) * 10 * KC * 2) % 128 == 0) hsplit = v;
  }
  a.hsplit = hsplit;
  B200_TRY(make_act_map(&maps.a[0], A.hi, A.N, A.D, A.H, A.W, A.C, A.ld, KC, 10, 18 / hsplit, 1, 1, swz_for_bytes(KC * 2), A.vD, A.vH, A.vW));
  B200_TRY(make_act_map(&maps.dy[0], Y.hi, Y.N, Y.D, Y.H, Y.W, Y.C, Y.ld, CBN, 8, 16, 1, 1, swz_for_bytes(CBN * 2), Y.vD, Y.vH, Y.vW));
  if (split) {
    B200_TRY(make_act_map(&maps.a[1], A.lo, A.N, A.D, A.H, A.W, A.C, A.ld, KC, 10, 18 / hsplit, 1, 1, swz_for_bytes(KC * 2), A.vD, A.vH, A.vW));
    B200_TRY(make_act_map(&maps.dy[1], Y.lo, Y.N, Y.D, Y.H, Y.W, Y.C, Y.ld, CBN, 8, 16, 1, 1, swz_for_bytes(CBN * 2), Y.vD, Y.vH, Y.vW));
  }
  dim3 grid((unsigned)splits, (unsigned)(a.nkc * a.gpk), (unsigned)cotiles);
#define B200_WGH_CASE(kc, bn) \
  if (KC == kc && BN == bn && TD == 2) return launch_wgh<kc, bn, 2>(maps, a, grid, st); \
  if (KC == kc && BN == bn && TD == 4 && bn <= 32) return launch_wgh<kc, (bn <= 32 ? bn : 32), 4>(maps, a, grid, st);
  B200_WGH_CASE(16, 16) B200_WGH_CASE(16, 32) B200_WGH_CASE(16, 64)
  B200_WGH_CASE(32, 16) B200_WGH_CASE(32, 32) B200_WGH_CASE(32, 64)
#undef B200_WGH_CASE
  set_error("wgrad_halo: no kernel for KC=%d BN=%d", KC, BN);
  return E_UNSUPPORTED;
}

}  // namespace b200
