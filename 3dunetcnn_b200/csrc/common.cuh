// Shared host/device declarations for libb200unet (internal; the public C ABI is include/b200unet.h).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace b200 {

typedef __nv_bfloat16 bf16;

// Status codes returned through the C ABI.
enum Status : int { OK = 0, E_INVALID = -1, E_UNSUPPORTED = -2, E_CUDA = -3, E_DRIVER = -4 };

void set_error(const char* fmt, ...);
const char* get_error();

#define B200_CHECK_CUDA(expr)                                                                 \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      b200::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return b200::E_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define B200_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      b200::set_error(__VA_ARGS__);    \
      return (code);                   \
    }                                  \
  } while (0)

#define B200_TRY(expr)        \
  do {                        \
    int _s = (expr);          \
    if (_s != b200::OK) return _s; \
  } while (0)

// NDHWC bf16 tensor view.  value = hi (+ lo when lo != nullptr: split-precision "parity" mode).
struct Act {
  bf16* hi;
  bf16* lo;
  int N, D, H, W, C;  // logical extent; C = channels visible through this view
  int ld;             // channel pitch of the underlying buffer in elements (>= C)
  int vD, vH, vW;     // optional "visible" spatial extents (0 = D/H/W): voxels beyond them read as zero through TMA
  __host__ __device__ long long voxels() const { return (long long)N * D * H * W; }
};

static inline Act make_act(bf16* hi, bf16* lo, int N, int D, int H, int W, int C, int ld) {
  Act a; a.hi = hi; a.lo = lo; a.N = N; a.D = D; a.H = H; a.W = W; a.C = C; a.ld = ld; a.vD = a.vH = a.vW = 0; return a;
}
static inline Act slice_c(const Act& a, int c0, int c) {
  Act r = a; r.hi = a.hi + c0; r.lo = a.lo ? a.lo + c0 : nullptr; r.C = c; return r;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace b200
