// warp_dsl.h -- a tiny "one warp = one sample" programming layer.
//
// The rollout core (rollout_core.h) is written once against these macros and compiled twice:
//   * by nvcc for sm_100a, where a LANES block is the body executed by each of the 32 lanes of the
//     warp that owns the sample, per-lane state lives in registers, collectives are warp shuffles
//     and every block is fenced by __syncwarp();
//   * by g++ with -DCEMK_EMU (tests/emu, no-GPU unit tests only), where a LANES block is a plain
//     `for (lane = 0..31)` loop over an array of per-lane register structs.  This is a debugging
//     aid for kernel logic in a container without a GPU; it is never used by the product path.
//
// Rules the core follows so both builds mean the same thing:
//   1. inside one LANES block a lane reads only shared scratch written in *earlier* blocks (or by
//      itself) and writes only locations no other lane touches in that block;
//   2. code outside LANES blocks is warp-uniform (every lane computes the same scalars); shared
//      scratch is written there only through UNIFORM_WRITE, which is fenced on both sides.
#pragma once
#include <math.h>

#ifdef CEMK_EMU
#include <cstring>
#define KFN static inline
#define KNOINLINE static
#define STEP_ALIGN()
#define PHASE_ALIGN(bit)
#define LANES(W, R) for (int lane = 0; lane < 32; ++lane) { auto& R = (W).regs[lane]; (void)R;
#define END_LANES }
#define RLANES(W, R) LANES(W, R)
#define END_RLANES }
#define UNIFORM_WRITE(W) if (true)
#define END_UNIFORM_WRITE
#define USYNC()
#define KRSQRT(x) (1.0f / sqrtf(x))
#define KPOPC(x) __builtin_popcount(x)
#define KFFS(x) __builtin_ffs((int)(x))
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
#else
#define KFN __device__ __forceinline__
#define KNOINLINE __device__ __noinline__
#ifdef CEMK_STEP_SYNC
// keep the live warps of a CTA on the same code (instruction-cache locality).  Named barrier 1 with the
// live thread count of this CTA (a CTA may run fewer samples than its launch width, see cemk_rollout_cost).
#define CTA_ALIGN(W) asm volatile("bar.sync 1, %0;" :: "r"((W).nthr) : "memory")
#define STEP_ALIGN() CTA_ALIGN(W)
#else
#define STEP_ALIGN()
#endif
// optional extra CTA-wide re-alignments inside a step, selected by the bits of CEMK_PHASE_SYNC
// (only at points every warp reaches)
#if defined(CEMK_STEP_SYNC) && defined(CEMK_PHASE_SYNC)
#define PHASE_ALIGN(bit) do { if ((CEMK_PHASE_SYNC) & (bit)) CTA_ALIGN(W); } while (0)
#else
#define PHASE_ALIGN(bit)
#endif
#define LANES(W, R) { __syncwarp(); const int lane = (W).lane; auto& R = (W).regs; (void)R;
#define END_LANES } __syncwarp();
// register-only lane block: touches no shared scratch, so no fences
#define RLANES(W, R) { const int lane = (W).lane; auto& R = (W).regs; (void)R; (void)lane;
#define END_RLANES }
#define UNIFORM_WRITE(W) __syncwarp(); if ((W).lane == 0)
#define END_UNIFORM_WRITE __syncwarp();
#define USYNC() __syncwarp()
#define KRSQRT(x) rsqrtf(x)
#define KPOPC(x) __popc(x)
#define KFFS(x) __ffs((int)(x))
#endif

template <class LR>
struct WarpCtx {
#ifdef CEMK_EMU
  LR regs[32];
#else
  LR regs;
  int lane;
  int nthr;          // live threads of this CTA (barrier width)
#ifdef CEMK_PHASE_TIMING
  long long t0; int phase; long long ph[24];
#endif
#endif
};

// Debug build only (-DCEMK_PHASE_TIMING): per-phase SM-clock accounting of one warp's step, summed
// into a global table by the kernel wrapper (tools/phase_timing.py).  No-op otherwise.
#if !defined(CEMK_EMU) && defined(CEMK_PHASE_TIMING)
#define PHASE(W, id) do { long long now_ = clock64(); (W).ph[(W).phase] += now_ - (W).t0; (W).t0 = now_; (W).phase = (id); } while (0)
#else
#define PHASE(W, id) do { } while (0)
#endif

// sum over lanes of f(lane, regs); result is warp-uniform
template <class W, class F>
KFN float warp_sum(W& w, F f) {
#ifdef CEMK_EMU
  float s = 0.f;
  for (int l = 0; l < 32; ++l) s += f(l, w.regs[l]);
  return s;
#else
  float v = f(w.lane, w.regs);
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
#endif
}

// exclusive prefix sum of small non-negative ints; set(lane, regs, offset) receives each lane's
// offset; returns the warp total
template <class W, class G, class S>
KFN int warp_excl_scan(W& w, G get, S set) {
#ifdef CEMK_EMU
  int run = 0;
  for (int l = 0; l < 32; ++l) { int v = get(l, w.regs[l]); set(l, w.regs[l], run); run += v; }
  return run;
#else
  int v = get(w.lane, w.regs), inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (w.lane >= o) inc += t; }
  set(w.lane, w.regs, inc - v);
  return __shfl_sync(0xffffffffu, inc, 31);
#endif
}

// value of f(src, regs[src]) broadcast to every lane
template <class W, class F>
KFN float warp_bcast(W& w, int src, F f) {
#ifdef CEMK_EMU
  return f(src, w.regs[src]);
#else
  return __shfl_sync(0xffffffffu, f(w.lane, w.regs), src);
#endif
}

// lane index of the maximum of f(lane, regs) over all lanes; ties -> lowest lane ("first max");
// lanes whose value is NaN never win against a number
template <class W, class F>
KFN int warp_argmax_first(W& w, F f) {
#ifdef CEMK_EMU
  int best = 0; float bv = f(0, w.regs[0]);
  for (int l = 1; l < 32; ++l) { float v = f(l, w.regs[l]); if (v > bv || (bv != bv && v == v)) { bv = v; best = l; } }
  return best;
#else
  float v = f(w.lane, w.regs); int idx = w.lane;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx) || (v != v && ov == ov)) { v = ov; idx = oi; }
  }
  return idx;
#endif
}

// per-lane-source shuffle: lane l receives get(src(l), regs[src(l)]) through set(l, regs[l], value)
template <class W, class G, class SRC, class SET>
KFN void warp_shfl_each(W& w, G get, SRC src, SET set) {
#ifdef CEMK_EMU
  float vals[32];
  for (int l = 0; l < 32; ++l) vals[l] = get(l, w.regs[l]);
  for (int l = 0; l < 32; ++l) set(l, w.regs[l], vals[src(l) & 31]);
#else
  const float v = __shfl_sync(0xffffffffu, get(w.lane, w.regs), src(w.lane));
  set(w.lane, w.regs, v);
#endif
}

// bit l of the result is set iff pred(l, regs[l]) holds
template <class W, class F>
KFN unsigned warp_ballot(W& w, F pred) {
#ifdef CEMK_EMU
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) if (pred(l, w.regs[l])) m |= 1u << l;
  return m;
#else
  return __ballot_sync(0xffffffffu, pred(w.lane, w.regs));
#endif
}
// warp_argmax_first restricted to lanes 0..7 (three exchange rounds); other lanes' values are ignored
template <class W, class F>
KFN int warp_argmax_first8(W& w, F f) {
#ifdef CEMK_EMU
  int best = 0; float bv = f(0, w.regs[0]);
  for (int l = 1; l < 8; ++l) { float v = f(l, w.regs[l]); if (v > bv || (bv != bv && v == v)) { bv = v; best = l; } }
  return best;
#else
  float v = f(w.lane, w.regs); int idx = w.lane;
#pragma unroll
  for (int o = 4; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx) || (v != v && ov == ov)) { v = ov; idx = oi; }
  }
  return __shfl_sync(0xffffffffu, idx, 0);
#endif
}

// Sums of nine per-lane values (regs.acc[0..8]) over the warp, all nine results to every lane.
// GPU: transposed butterfly -- each exchange round halves the number of values a lane still carries
// (5+3+2+1+1 shuffles), then nine broadcasts: 21 shuffles instead of 45.
template <class W>
KFN void warp_sum9(W& w, float* out) {
#ifdef CEMK_EMU
  for (int k = 0; k < 9; ++k) { float s = 0.f; for (int l = 0; l < 32; ++l) s += w.regs[l].acc[k]; out[k] = s; }
#else
  const int lane = w.lane;
  const bool hi = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
  float v[10];
#pragma unroll
  for (int k = 0; k < 9; ++k) v[k] = w.regs.acc[k];
  v[9] = 0.f;
  float u[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? v[k] : v[5 + k], 16);
    u[k] = (hi ? v[5 + k] : v[k]) + recv;
  }
  float t[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float mine = h8 ? (k < 2 ? u[3 + k] : 0.f) : u[k];
    const float send = h8 ? u[k] : (k < 2 ? u[3 + k] : 0.f);
    t[k] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  float s2[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float mine = h4 ? (k < 1 ? t[2] : 0.f) : t[k];
    const float send = h4 ? t[k] : (k < 1 ? t[2] : 0.f);
    s2[k] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  float r = (h2 ? s2[1] : s2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? s2[0] : s2[1], 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  // lane that ends up holding the total of original index k
  out[0] = __shfl_sync(0xffffffffu, r, 0);  out[1] = __shfl_sync(0xffffffffu, r, 2);  out[2] = __shfl_sync(0xffffffffu, r, 4);
  out[3] = __shfl_sync(0xffffffffu, r, 8);  out[4] = __shfl_sync(0xffffffffu, r, 10); out[5] = __shfl_sync(0xffffffffu, r, 16);
  out[6] = __shfl_sync(0xffffffffu, r, 18); out[7] = __shfl_sync(0xffffffffu, r, 20); out[8] = __shfl_sync(0xffffffffu, r, 24);
#endif
}
