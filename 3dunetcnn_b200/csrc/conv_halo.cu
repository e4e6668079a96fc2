// Halo-resident implicit-GEMM convolution (3x3x3 stride 1, optional fused 1x1x1 second source, or a lone 1x1x1) for
// sm_100a: the dominant kernel of the training step.  Used whenever the output plane is at least 8 x 16 and either
// Cin <= 64 or there are enough tiles (see conv_halo_eligible); igemm_conv.cu's per-tap streaming kernel takes the rest.
//
// A CTA (persistent, one per SM) walks output tiles of 8(w) x 16(h) x TD(d) voxels x BN channels.  For each K chunk of
// KC input channels it TMA-loads ONE halo box (KC, 10, 18, TD+2) of the activation into shared memory and serves all 27
// taps from it: tap (kd,kh,kw) of output plane d is the UMMA A operand whose descriptor START ADDRESS is shifted by
// ((d+kd)*18 + kh)*10 + kw rows and whose 8-row-group stride (SBO) is one halo row of 10 voxels.  The hardware applies
// the 64B/32B swizzle on absolute shared-memory address bits (verified by csrc/probe.cu on B200, profiles/probe_r01.txt),
// so shifted / re-strided descriptors read exactly what TMA wrote.  L2->SM activation traffic drops from 27 reads per
// voxel (streaming kernel) to (TD+2)/TD * 1.41.
// Weights stream through a ring of stages; for BN <= 64 a stage holds the kd = 0,1,2 tiles of one (kh,kw) and the MMAs
// are stacked along N (one MMA of N = 3*BN feeds three output planes: see the issuer).  TD accumulators of 128 x BN fp32
// live in TMEM, in NACC sets so the epilogue of tile i overlaps the MMAs of tile i+1; two halo buffers let the next
// chunk / tile load under the current MMAs.
// Warps: 0 = weight-stage producer, 10 = halo producer, 1 = MMA issuer (one elected lane per stage), 2..9 = epilogue in
// two groups of four.  Epilogue: TMEM -> registers -> (+residual)(*dropout)(+bias) | GroupNorm/ReLU backward -> swizzled
// shared staging tile -> TMA store (full-line writes; the per-thread 16-byte global stores of the first version cost
// ~40 cycles per touched line and bounded every narrow layer).  Per-channel statistics accumulate in registers across
// planes and tiles (BN <= 64) and are reduced by a transposing warp butterfly only when the sample changes; their
// reduction order is fixed, so repeated launches give bit-identical statistics up to the final fp64 atomics.
#include "conv_halo_kernel.cuh"

namespace b200 {

// the lean epilogue variants live in their own translation units (parallel compilation)
int launch_halo_ev1(int KC, int BN, int TD, int kws, const ConvMaps& maps, const ConvArgs& a, const HaloArgs& h, int grid, cudaStream_t st);
int launch_halo_ev2(int KC, int BN, int TD, int kws, const ConvMaps& maps, const ConvArgs& a, const HaloArgs& h, int grid, cudaStream_t st);

bool conv_halo_eligible(const ConvOp& op) {
  // measured on B200 (profiles/r01_final_convbench.jsonl): the halo-resident kernel always wins for Cin <= 64.  For wider
  // inputs both kernels are bound by L2 -> SM operand traffic (the streaming kernel re-reads the activation 27 times);
  // the halo kernel with 64-channel output tiles and kd-stacked N = 192 MMAs needs about a third of it and wins (1470 vs
  // 1100 TFLOP/s at 128ch@64^3) once there are enough tiles to fill the SMs: voxels * Cout >= 2^24.
  static const int max_c = getenv("B200UNET_HALO_MAXC") ? atoi(getenv("B200UNET_HALO_MAXC")) : 512;
  if (op.src[0].x.C > max_c) return false;
  long long wide_min = 1LL << 24;   // tests lower it (B200UNET_HALO_WIDE_MIN=0) to reach this path with small oracle-sized shapes
  if (const char* e = getenv("B200UNET_HALO_WIDE_MIN")) wide_min = atoll(e);
  if (op.src[0].x.C > 64 && (long long)op.out.N * op.out.D * op.out.H * op.out.W * op.out.C < wide_min) return false;
  // a lone 1x1x1 convolution (the residual blocks' `sample` data gradient) runs as the kernel's centre-tap source: two
  // MMAs per streaming-kernel tile left that launch bound by per-CTA set-up (0.75 ms for 32->64 at 128^3)
  static const bool no_1x1 = getenv("B200UNET_NO_HALO_1X1") != nullptr;
  if (op.nsrc == 1 && op.src[0].ksz == 1 && op.src[0].stride == 1) return !no_1x1 && op.out.W >= 8 && op.out.H >= 16;
  if (op.src[0].ksz != 3 || op.src[0].stride != 1) return false;
  if (op.nsrc == 2 && (op.src[1].ksz != 1 || op.src[1].stride != 1)) return false;
  return op.out.W >= 8 && op.out.H >= 16;
}

// does (KC, BN, TD, split) fit the shared-memory budget with at least a 2-deep weight ring and 1 staging buffer?
static bool halo_fits(int KC, int BN, int TD, bool split, int kws = 1) {
  const int halo = (180 * (TD + 2) * KC * 2 + 1023) / 1024 * 1024;
  const int tpb = BN <= 64 ? 3 : 1;
  const int box = tpb * BN * KC * 2 * kws;
  const int bbytes = (box + 1023) / 1024 * 1024;
  const int aux = 1024 + 8 * BN * 8 + BN * 16;
  const int out_buf = 128 * BN * 2 * (split ? 2 : 1);
  return 232448 - 1024 - aux - 2 * halo - 2 * out_buf >= 2 * bbytes;
}

int launch_conv_halo(const ConvOp& op_in, int num_sms, cudaStream_t st) {
  B200_REQUIRE(conv_halo_eligible(op_in), E_UNSUPPORTED, "conv_halo: shape not eligible");
  ConvOp op = op_in;
  const bool only_1x1 = op.nsrc == 1 && op.src[0].ksz == 1;
  if (only_1x1) {   // centre-tap source lives in slot 1; slot 0 (the 27-tap source) stays empty
    op.src[1] = op.src[0];
    op.nsrc = 2;
  }
  const int s_begin = only_1x1 ? 1 : 0;
  const Act& out = op.out;
  B200_REQUIRE(out.C % 8 == 0 && out.ld % 8 == 0, E_UNSUPPORTED, "conv_halo: Cout=%d must be a multiple of 8", out.C);
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  a.N = out.N; a.Do = out.D; a.Ho = out.H; a.Wo = out.W; a.Cout = out.C;
  int cin_max = 0;
  bool split = false;
  for (int s = s_begin; s < op.nsrc; ++s) {
    const ConvSrc& c = op.src[s];
    B200_REQUIRE(c.x.C % 8 == 0 && c.x.ld % 8 == 0, E_UNSUPPORTED, "conv_halo: Cin=%d must be a multiple of 8", c.x.C);
    B200_REQUIRE(c.x.N == out.N && c.x.D == out.D && c.x.H == out.H && c.x.W == out.W, E_INVALID,
                 "conv_halo: source %d extent mismatch", s);
    if (c.x.C > cin_max) cin_max = c.x.C;
    if (c.x.lo || c.w_lo) split = true;
  }
  if (split) {
    for (int s = s_begin; s < op.nsrc; ++s)
      B200_REQUIRE(op.src[s].x.lo && op.src[s].w_lo, E_INVALID, "conv_halo: split mode needs lo parts on every source");
    B200_REQUIRE(out.lo != nullptr, E_INVALID, "conv_halo: split mode needs a lo output");
  }
  const int KC = cin_max > 16 ? 32 : 16;
  int BN = out.C > 64 ? 128 : out.C > 32 ? 64 : out.C > 16 ? 32 : 16;
  if (cin_max > 64 && BN > 64) BN = 64;   // wide inputs: 64-channel tiles so that the kd taps stack to N = 192
  if (const char* e = getenv("B200UNET_HALO_BN")) {   // tuning override: cap the N tile
    const int v = atoi(e);
    if ((v == 16 || v == 32 || v == 64 || v == 128) && v < BN) BN = v;
  }
  const int ntiles = ceil_div(out.C, BN);
  // TD: deepest tile that still gives every SM work (>= ~2 tiles per SM), bounded by TMEM (TD*BN <= 512) and smem
  int TD = 4;
  auto tiles_for = [&](int td) { return (long long)out.N * ceil_div(out.D, td) * ceil_div(out.H, 16) * ceil_div(out.W, 8) * ntiles; };
  while (TD > 1 && (tiles_for(TD) < 2LL * num_sms || out.D < TD || !halo_fits(KC, BN, TD, split))) TD >>= 1;
  if (const char* e = getenv("B200UNET_HALO_TD")) {   // tuning override (1, 2 or 4)
    const int v = atoi(e);
    if ((v == 1 || v == 2 || v == 4) && v * BN <= 512 && halo_fits(KC, BN, v, split)) TD = v;
  }
  B200_REQUIRE(halo_fits(KC, BN, TD, split), E_UNSUPPORTED, "conv_halo: KC=%d BN=%d does not fit shared memory", KC, BN);
  a.tw = 8; a.th = 16; a.td = TD;
  a.tiles_w = ceil_div(out.W, 8); a.tiles_h = ceil_div(out.H, 16); a.tiles_d = ceil_div(out.D, TD);
  const Swz swz = swz_for_bytes(KC * 2);
  // sub-boxes must start on 128-byte boundaries: rows * RB % 128 == 0
  int hsplit = KC == 32 ? 2 : 1;
  if (const char* e = getenv("B200UNET_HALO_HSPLIT")) {
    const int v = atoi(e);
    if ((v == 1 || v == 2 || v == 3 || v == 6) && ((18 / v) * 10 * KC * 2) % 128 == 0) hsplit = v;
  }
  const int tpb = BN <= 64 ? 3 : 1;
  static const bool dense1 = !(getenv("B200UNET_HALO_1X1_DENSE") && atoi(getenv("B200UNET_HALO_1X1_DENSE")) == 0);   // A/B switch
  for (int s = s_begin; s < op.nsrc; ++s) {
    const ConvSrc& c = op.src[s];
    a.ntaps[s] = c.ksz * c.ksz * c.ksz; a.ksz[s] = c.ksz; a.stride[s] = 1;
    a.kchunks[s] = ceil_div(c.x.C, KC);
    const int boxT = (s == 0) ? tpb : 1;
    // source 0 (27 taps): halo boxes (KC, 10, 18/hsplit, 1); source 1 (centre tap only): one dense box (KC, 8, 16, TD)
    const bool dense = dense1 && s == 1;
    const int bxw = dense ? 8 : 10, bxh = dense ? 16 : 18 / hsplit, bxd = dense ? TD : 1;
    B200_TRY(make_act_map(&maps.a[s][0], c.x.hi, c.x.N, c.x.D, c.x.H, c.x.W, c.x.C, c.x.ld, KC, bxw, bxh, bxd, 1, swz,
                          c.x.vD, c.x.vH, c.x.vW));
    const bool stk = (s == 0) && BN <= 64;   // kd-stacked weight stages: 4-D view (Cin, Cout, khkw, kd), box (KC, BN, 1, 3)
    if (stk) B200_TRY(make_w_map_kd(&maps.b[s][0], c.w_hi, op.Cop, c.Cip, KC, BN, swz));
    else B200_TRY(make_w_map(&maps.b[s][0], c.w_hi, a.ntaps[s], op.Cop, c.Cip, KC, BN, swz, boxT));
    if (split) {
      B200_TRY(make_act_map(&maps.a[s][1], c.x.lo, c.x.N, c.x.D, c.x.H, c.x.W, c.x.C, c.x.ld, KC, bxw, bxh, bxd, 1, swz,
                            c.x.vD, c.x.vH, c.x.vW));
      if (stk) B200_TRY(make_w_map_kd(&maps.b[s][1], c.w_lo, op.Cop, c.Cip, KC, BN, swz));
      else B200_TRY(make_w_map(&maps.b[s][1], c.w_lo, a.ntaps[s], op.Cop, c.Cip, KC, BN, swz, boxT));
    }
  }
  if (only_1x1) {   // slot 0 is never loaded (kchunks[0] = 0); keep its descriptors valid for the prefetch
    maps.a[0][0] = maps.a[1][0];
    maps.b[0][0] = maps.b[1][0];
  }
  // output tile stores: box (min(BN,64) channels, 8, 16, 1, 1), swizzle by the box row bytes
  const int cbo = BN < 64 ? BN : 64;
  B200_TRY(make_act_map(&maps.o[0], out.hi, out.N, out.D, out.H, out.W, out.C, out.ld, cbo, 8, 16, 1, 1, swz_for_bytes(cbo * 2)));
  if (split)
    B200_TRY(make_act_map(&maps.o[1], out.lo, out.N, out.D, out.H, out.W, out.C, out.ld, cbo, 8, 16, 1, 1, swz_for_bytes(cbo * 2)));
  a.npass = split ? 3 : 1;
  a.mode = op.mode;
  a.out_hi = out.hi; a.out_lo = out.lo; a.ldo = out.ld;
  if (op.res) {
    B200_REQUIRE(op.res->C == out.C, E_INVALID, "conv_halo: residual channel mismatch");
    a.res_hi = op.res->hi; a.res_lo = op.res->lo; a.ldr = op.res->ld;
  }
  a.scale = op.scale;
  a.bias = op.bias; a.zero_last = op.zero_last;
  a.stats = op.stats; a.stats_ld = op.stats_ld;
  if (op.mode == 1) {
    B200_REQUIRE(op.gn_x && op.coef, E_INVALID, "conv_halo: mode 1 needs gn_x and coef");
    B200_REQUIRE(op.gn_x->C == out.C, E_INVALID, "conv_halo: gn_x channel mismatch");
    a.x_hi = op.gn_x->hi; a.x_lo = op.gn_x->lo; a.ldx = op.gn_x->ld;
    a.coef = reinterpret_cast<const float4*>(op.coef); a.coef_ld = op.coef_ld;
    a.slope = op.slope; a.bstats = op.bstats;
  }
  HaloArgs h;
  memset(&h, 0, sizeof(h));
  h.ntiles = ntiles;
  h.hsplit = hsplit;
  h.split = split ? 1 : 0;
  h.fd_nt = make_fastdiv((unsigned)ntiles); h.fd_w = make_fastdiv((unsigned)a.tiles_w);
  h.fd_h = make_fastdiv((unsigned)a.tiles_h); h.fd_d = make_fastdiv((unsigned)a.tiles_d);
  h.dense1 = dense1 ? 1 : 0;
  static const bool prefetch = getenv("B200UNET_HALO_PREFETCH") && atoi(getenv("B200UNET_HALO_PREFETCH")) == 1;
  h.prefetch = prefetch ? 1 : 0;
  // side input of the epilogue (residual / GroupNorm input): same box and swizzle as the output tile stores, so that the lean
  // epilogues can TMA-load it straight into the output staging buffers (HaloArgs::side_ring)
  static const bool no_ring = getenv("B200UNET_HALO_SIDE_RING") && atoi(getenv("B200UNET_HALO_SIDE_RING")) == 0;   // A/B switch
  const Act* side = op.mode == 1 ? op.gn_x : op.res;
  maps.side = maps.o[0];
  if (side && side->hi)
    B200_TRY(make_act_map(&maps.side, side->hi, side->N, side->D, side->H, side->W, side->C, side->ld, cbo, 8, 16, 1, 1,
                          swz_for_bytes(cbo * 2)));
  h.dbg = nullptr;
  if (const char* e = getenv("B200UNET_HALO_DBG")) h.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  h.tiles_total = (int)tiles_for(TD);
  const int grid = h.tiles_total < num_sms ? h.tiles_total : num_sms;
  // Three-box weight stages (KW = 3) where they measured faster on B200 (profiles/r02_convbench*.log, batch-2 128^3 layers):
  //   64->32 all epilogues -15 %, 32->32 with a plain / statistics epilogue -10 %, 32->8 GroupNorm-backward epilogue -17 %;
  //   but 32->32 with a side input (residual, GroupNorm-backward) +12..25 % and 8->32 +6 % -> those keep one box per stage.
  // (The round-1 experiment with two issuing warps and private accumulator sets, NI = 2, ran correctly but slower than
  //  either variant -- 0.38 vs 0.30 ms on 32->32+residual -- and is no longer instantiated.)
  const bool side_input = op.res != nullptr || op.mode == 1;
  // epilogue variant: the lean bodies when none of the rare options is in play
  static const bool no_lean = getenv("B200UNET_HALO_GENERIC_EPILOGUE") != nullptr;   // A/B switch
  const bool lean = !no_lean && !split && !op.scale && !op.bias && !op.zero_last && !(no_ring && side && side->hi);
  int kws = 1;
  if (BN <= 32 && (tpb * BN * KC * 2) % 1024 == 0 && KC == 32) {
    // with the side ring (lean variants) the three-box stages also win on side-input layers: 32->32 GroupNorm-backward
    // 0.285 -> 0.246 ms (profiles/r02_layer_times_kws3_ab.txt); the generic variant's register-staged side rows keep the old rule
    if (a.kchunks[0] >= 2 || BN < 32 || !side_input || lean) kws = 3;
  }
  if (const char* e = getenv("B200UNET_HALO_KWS")) {   // tuning override (1 or 3)
    const int v = atoi(e);
    if (v == 1 || (v == 3 && BN <= 32 && (tpb * BN * KC * 2) % 1024 == 0)) kws = v;
  }
  if (kws == 3 && !halo_fits(KC, BN, TD, split, 3)) kws = 1;
  if (lean) {
    h.side_ring = (side && side->hi) ? 1 : 0;
    int rc = op.mode == 0 ? launch_halo_ev1(KC, BN, TD, kws, maps, a, h, grid, st) : launch_halo_ev2(KC, BN, TD, kws, maps, a, h, grid, st);
    if (rc == HALO_NO_RING && kws == 3)   // three-box weight stages leave no room: one box per stage with the ring
      rc = op.mode == 0 ? launch_halo_ev1(KC, BN, TD, 1, maps, a, h, grid, st) : launch_halo_ev2(KC, BN, TD, 1, maps, a, h, grid, st);
    if (rc != HALO_NO_RING) return rc;
    h.side_ring = 0;     // three staging buffers per group do not fit: generic variant, side rows through registers
  }
  return launch_halo_table<0>(KC, BN, TD, kws, maps, a, h, grid, st);
}

}  // namespace b200
