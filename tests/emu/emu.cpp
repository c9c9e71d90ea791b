// tests/emu/emu.cpp -- CPU single-stepping harness for the rollout core (NO-GPU UNIT TESTS ONLY).
//
// Compiles manipulator_mujoco_b200/csrc/rollout_core.h with -DCEMK_EMU, where each LANES block
// becomes a loop over KW per-lane register structs (see warp_dsl.h).  It exists so the kernel's
// indexing / physics logic can be checked against the oracle in a container without a GPU.
// Nothing in the package imports it; the product path is the CUDA build in csrc/cemk.cu.
#define CEMK_EMU 1
#include "../../manipulator_mujoco_b200/csrc/rollout_core.h"
#include <cstdlib>
#include <new>

extern "C" int emu_sizeof_kmodel() { return (int)sizeof(KModel); }

// Race-checking build (-DCEMK_EMU_RACE, see warp_dsl.h): counters over all rollouts since the last read.
static long long g_race_waw = 0, g_race_blocks = 0;
static int g_race_first[3] = {0, 0, 0};
extern "C" int emu_race_enabled() {
#ifdef CEMK_EMU_RACE
  return 1;
#else
  return 0;
#endif
}
extern "C" void emu_race_stats(long long* out5) {
  out5[0] = g_race_waw; out5[1] = g_race_blocks; out5[2] = g_race_first[0]; out5[3] = g_race_first[1]; out5[4] = g_race_first[2];
  g_race_waw = g_race_blocks = 0; g_race_first[0] = 0;
}

// Deliberately broken lane blocks, to show that the checker sees what it is meant to see (tests/test_emu_race.py).
//   which = 1: a lane reads what its neighbour wrote in the same block (missing fence, rule 1);
//   which = 2: two lanes write different values to the same scratch word in one block;
//   which = 0: the correctly fenced version of 1 (write block, then read block).
// out[0..KW-1] = what each lane read, out[KW] = write-write conflicts counted.
extern "C" void emu_race_selftest(int which, float* out) {
  Warp* W = new Warp();
  WarpSmemT<KM_NC_FAST>* S = new WarpSmemT<KM_NC_FAST>();
  memset((void*)W, 0, sizeof(Warp));
  memset((void*)S, 0, sizeof(*S));
  long long waw = 0;
#ifdef CEMK_EMU_RACE
  RaceState rs;
  rs.add(S, sizeof(*S));
  W->race = &rs;
#endif
  Warp& Wr = *W;
  if (which == 1) {
    LANES(Wr, R)
      if (lane < KM_NV) S->qacc[lane] = (float)(lane + 1);
      R.f0 = (lane > 0 && lane <= KM_NV) ? S->qacc[lane - 1] : 0.f;      // BUG on purpose: same-block read of another lane's write
    END_LANES
  } else if (which == 2) {
    LANES(Wr, R)
      if (lane < 2) S->qacc[0] = (float)(lane + 1);                       // BUG on purpose: two writers, different values
      R.f0 = 0.f;
    END_LANES
  } else {
    LANES(Wr, R)
      if (lane < KM_NV) S->qacc[lane] = (float)(lane + 1);
    END_LANES
    LANES(Wr, R)
      R.f0 = (lane > 0 && lane <= KM_NV) ? S->qacc[lane - 1] : 0.f;
    END_LANES
  }
#ifdef CEMK_EMU_RACE
  waw = rs.waw;
#endif
  for (int l = 0; l < KW; ++l) out[l] = W->regs[l].f0;
  out[KW] = (float)waw;
  delete W; delete S;
}

extern "C" int emu_rollout(const KModel* m, int B, int T, const float* thetadot, const float* q0, const float* v0,
                           const float* tpos, const float* trot, float w_pos, float w_rot, float w_col,
                           float* theta, float* cost4, float* eef_pos, float* eef_rot, float* collision,
                           float* qacc_dbg, int* flags, int nc) {
#pragma omp parallel
  {
    Warp* W = new Warp();
    WarpSmemT<KM_NC_FAST>* S = new WarpSmemT<KM_NC_FAST>();
    WarpSmemT<KM_NC_BIG>* Sb = new WarpSmemT<KM_NC_BIG>();
    float* prevd = new float[2 * KM_NPASS * KW];
    float* ovf = new float[KM_NC_TOT * KM_OVF_STRIDE];
#pragma omp for schedule(dynamic, 4)
    for (int s = 0; s < B; ++s) {
      memset((void*)W, 0, sizeof(Warp));
      memset((void*)S, 0, sizeof(*S));
      memset((void*)Sb, 0, sizeof(*Sb));
      RolloutArgs A;
      A.T = T;
      A.live = true;
      A.thetadot = thetadot + (size_t)s * KM_NL * T;
      A.q0 = q0; A.v0 = v0; A.target_pos = tpos; A.target_rot = trot;
      A.w_pos = w_pos; A.w_rot = w_rot; A.w_col = w_col;
      A.theta = theta + (size_t)s * KM_NL * T;
      A.cost4 = cost4 + (size_t)s * 4;
      A.eef_pos = eef_pos ? eef_pos + (size_t)s * T * 3 : nullptr;
      A.eef_rot = eef_rot ? eef_rot + (size_t)s * T * 4 : nullptr;
      A.collision = collision ? collision + (size_t)s * T * m->nslot_robot : nullptr;
      A.qacc_dbg = qacc_dbg ? qacc_dbg + (size_t)s * T * KM_NV : nullptr;
      A.flags = flags ? flags + s : nullptr;
      A.prevd = prevd;
      A.ovf = ovf;
#ifdef CEMK_EMU_RACE
      RaceState rs;                                   // shared scratch of this sample: the record, the previous distances, the spill area
      if (nc <= KM_NC_FAST) rs.add(S, sizeof(*S)); else rs.add(Sb, sizeof(*Sb));
      rs.add(prevd, sizeof(float) * 2 * KM_NPASS * KW);
      rs.add(ovf, sizeof(float) * KM_NC_TOT * KM_OVF_STRIDE);
      W->race = &rs;
      g_emu_race = &rs;
#endif
      if (nc <= KM_NC_FAST) rollout_sample<KM_NC_FAST>(*W, *m, *S, A); else rollout_sample<KM_NC_BIG>(*W, *m, *Sb, A);
#ifdef CEMK_EMU_RACE
#pragma omp critical
      { g_race_waw += rs.waw; g_race_blocks += rs.blocks;
        if (rs.waw && !g_race_first[0]) { g_race_first[0] = rs.first_line; g_race_first[1] = rs.first_word; g_race_first[2] = rs.first_lanes; } }
      delete[] rs.snap; delete[] rs.pend; delete[] rs.owner;
#endif
    }
    delete W; delete S; delete Sb; delete[] prevd; delete[] ovf;
  }
  return 0;
}

// the kernel's capsule-box collider in isolation (full contact record), for fuzzing against the oracle's
extern "C" void emu_capsule_box(const float* A, const float* B, float r, const float* bpos, const float* bmat, const float* bsize,
                                float* dist2, float* pos6, float* nrm6) {
  Contact2 c;
  capsule_box<true>(A, B, r, bpos, bmat, bsize, c);
  for (int j = 0; j < 2; ++j) {
    dist2[j] = c.dist[j];
    for (int k = 0; k < 3; ++k) { pos6[3 * j + k] = c.pos[j][k]; nrm6[3 * j + k] = c.nrm[j][k]; }
  }
}

// The near pass spreads the edge stage of a pair over 12 lanes (capbox_edge_lane + first-maximum reduction); this is the same
// computation done serially, for fuzzing against capsule_box<true> (which runs capbox_edges): dist2 / pos6 / nrm6 in box coordinates.
extern "C" void emu_capsule_box_lanes(const float* a, const float* b, float r, const float* bsize, float* dist2, float* pos6, float* nrm6,
                                      float* dist2_serial, float* pos6_serial, float* nrm6_serial) {
  CapBoxOut c, cs;
  float nf[3];
  const int nout = capsule_box_near<true, false>(a, b, r, bsize, c, nf);
  capsule_box_near<true, true>(a, b, r, bsize, cs);
  if (nout >= 2) {
    float best = -INFINITY; EdgeEval be; be.epen = -1.f;
    for (int e = 0; e < 12; ++e) { const EdgeEval ev = capbox_edge_lane(e, a, b, r, bsize); if (ev.epen > best) { best = ev.epen; be = ev; } }
    const float minface = fminf(-c.dist[0], -c.dist[1]);
    if (best > 0.f) {
      const bool parallel = fabsf(be.dir[0] * nf[0] + be.dir[1] * nf[1] + be.dir[2] * nf[2]) > 0.99f;
      if ((minface > 0.f ? best < minface : true) && !parallel) {
        c.dist[0] = -best;
        for (int k = 0; k < 3; ++k) { c.pos[0][k] = 0.5f * (be.pa[k] + be.pb[k] + be.dir[k] * r); c.nrm[0][k] = be.dir[k]; }
      }
    }
  }
  for (int j = 0; j < 2; ++j) {
    dist2[j] = c.dist[j]; dist2_serial[j] = cs.dist[j];
    for (int k = 0; k < 3; ++k) { pos6[3 * j + k] = c.pos[j][k]; nrm6[3 * j + k] = c.nrm[j][k]; pos6_serial[3 * j + k] = cs.pos[j][k]; nrm6_serial[3 * j + k] = cs.nrm[j][k]; }
  }
}
