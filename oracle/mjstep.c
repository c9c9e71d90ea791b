/*
 * oracle/mjstep.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (scalar, one sample at a time, double precision by default) of the physics the
 * reference planner obtains from `mjx.step` inside `cem_planner.compute_rollout_single`
 * (reference sampling_based_planner/mjx_planner.py:251-274).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load this file; the product
 * path (manipulator_mujoco_b200/csrc) never does.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party packages
 * mujoco==3.3.1 / mujoco-mjx==3.3.1 (reference requirements.txt:16-17) which are neither vendored
 * in the reference nor installable here, and the reference has no tests or golden vectors for it.
 * Every function below restates the published MuJoCo / MJX algorithm (SURVEY.md appendix B and
 * B-detail); the section tags (B.2, BD.5, ...) refer to that document.  What *is* pinned:
 * forward kinematics against the tcp / hande poses of SURVEY.md section 4, geom ids against
 * view_traj_mjx.py:54, and smooth dynamics against data/theta.csv + data/thetadot.csv
 * (tests/test_oracle_pins.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).  -DORACLE_FLOAT gives a float32 build
 * used only as the timed CPU baseline.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_COUNT
#include "count_real.h"      /* C++ build: `real` tallies its arithmetic per stage (profiles/r2_flop_count.json) */
thread_local long long g_cnt[ORACLE_NSTAGE];
thread_local int g_stage = 7;
#elif defined(ORACLE_FLOAT)
typedef float real;
#define RSQRT sqrtf
#define RFABS fabsf
#define RSIN sinf
#define RCOS cosf
#define RPOW powf
#else
typedef double real;
#define RSQRT sqrt
#define RFABS fabs
#define RSIN sin
#define RCOS cos
#define RPOW pow
#endif

#ifndef ORACLE_COUNT
#define OSTAGE(k)
#define OINACTIVE_BEGIN(active)
#define OINACTIVE_END()
#else
/* Work MJX does for a contact slot that turns out inactive (position, frame) feeds nothing: the op counter books it
 * under stage 7 ("dense, unobservable") instead of the collider stage. */
#define OINACTIVE_BEGIN(active) int ostage_keep_ = g_stage; if (!(active)) g_stage = 7
#define OINACTIVE_END() g_stage = ostage_keep_
#endif

#define MAXB 24      /* bodies */
#define MAXJ 12      /* joints */
#define MAXV 16      /* dofs */
#define MAXQ 20      /* qpos */
#define MAXG 24      /* collision geoms */
#define MAXP 128     /* candidate pairs */
#define MAXCON 256   /* contact slots */
#define MAXEFC (MAXJ + 4 * MAXCON)

#define G_PLANE 0
#define G_CAPSULE 3
#define G_BOX 6
#define J_FREE 0
#define J_HINGE 3

#define MJ_MINVAL 1e-15
#define MJ_MINIMP 0.0001
#define MJ_MAXIMP 0.9999

/* Everything is passed as doubles / ints from python (ctypes.Structure mirror in oracle/oracle.py). */
typedef struct {
  int nq, nv, nbody, njnt, ngeom, npair, ncon;
  int iterations, ls_iterations;
  int tcp_body, hande_body;
  int capbox_mode;                  /* capsule_box restatement, see capsule_box(): 1 = has_support gate (default), 0 = round-1 */
  double timestep, tolerance, ls_tolerance, impratio, meaninertia;
  double gravity[3];
  double tcp_pos[3];
  int body_parent[MAXB], body_jnt[MAXB], body_rootid[MAXB], body_weldid[MAXB];
  double body_pos[MAXB][3], body_quat[MAXB][4], body_mass[MAXB], body_ipos[MAXB][3];
  double body_inertia[MAXB][9], body_gravcomp[MAXB], body_invweight0[MAXB];
  int jnt_type[MAXJ], jnt_body[MAXJ], jnt_qposadr[MAXJ], jnt_dofadr[MAXJ], jnt_limited[MAXJ];
  double jnt_axis[MAXJ][3], jnt_range[MAXJ][2], jnt_armature[MAXJ], jnt_damping[MAXJ];
  double jnt_solref[MAXJ][2], jnt_solimp[MAXJ][5], jnt_margin[MAXJ];
  double dof_invweight0[MAXV];
  int geom_type[MAXG], geom_body[MAXG];
  double geom_pos[MAXG][3], geom_quat[MAXG][4], geom_size[MAXG][3];
  double geom_friction[MAXG][3], geom_solref[MAXG][2], geom_solimp[MAXG][5];
  int pair_g1[MAXP], pair_g2[MAXP], pair_slotadr[MAXP], pair_nslot[MAXP];
  int slot_robot[MAXCON];           /* 1 if the slot's pair involves a robot_i geom (mjx_planner.py:115) */
} omodel;

/* per-sample working set */
typedef struct {
  real qpos[MAXQ], qvel[MAXV], qacc_warmstart[MAXV], qacc[MAXV];
  real xpos[MAXB][3], xquat[MAXB][4], xmat[MAXB][9], xipos[MAXB][3];
  real xanchor[MAXJ][3], xaxis[MAXJ][3];
  real gpos[MAXG][3], gmat[MAXG][9];
  real subtree_com[MAXB][3];
  real cinert[MAXB][10], crb[MAXB][10], cdof[MAXV][6], cdof_dot[MAXV][6], cvel[MAXB][6];
  int dof_body[MAXV], dof_parent[MAXV];
  real M[MAXV][MAXV], L[MAXV][MAXV];
  real qfrc_passive[MAXV], qfrc_bias[MAXV], qfrc_smooth[MAXV], qacc_smooth[MAXV];
  /* contacts */
  real con_dist[MAXCON], con_pos[MAXCON][3], con_frame[MAXCON][9];
  int con_b1[MAXCON], con_b2[MAXCON], con_g1[MAXCON], con_g2[MAXCON];
  /* constraints (active rows only; inactive MJX rows are exact zeros, see B.4) */
  int nefc;
  real efc_J[MAXEFC][MAXV], efc_D[MAXEFC], efc_aref[MAXEFC], efc_pos[MAXEFC];
  real site_tcp[3];
} odata;

/* ------------------------------------------------------------------ small vector helpers */
static inline real dot3(const real *a, const real *b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static inline void cross3(real *r, const real *a, const real *b) {
  real x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void sub3(real *r, const real *a, const real *b) { r[0]=a[0]-b[0]; r[1]=a[1]-b[1]; r[2]=a[2]-b[2]; }
static inline void add3(real *r, const real *a, const real *b) { r[0]=a[0]+b[0]; r[1]=a[1]+b[1]; r[2]=a[2]+b[2]; }
static inline void addscl3(real *r, const real *a, const real *b, real s) { r[0]=a[0]+s*b[0]; r[1]=a[1]+s*b[1]; r[2]=a[2]+s*b[2]; }
static inline void scl3(real *r, const real *a, real s) { r[0]=a[0]*s; r[1]=a[1]*s; r[2]=a[2]*s; }
static inline void copy3(real *r, const real *a) { r[0]=a[0]; r[1]=a[1]; r[2]=a[2]; }
static inline real norm3(const real *a) { return RSQRT(dot3(a, a)); }
/* MJX math.normalize_with_norm: zero vector stays zero */
static inline real normalize3(real *a) {
  real n = norm3(a);
  if (n > 0) { a[0] /= n; a[1] /= n; a[2] /= n; } else { a[0] = a[1] = a[2] = 0; }
  return n;
}
static inline void mat_vec(real *r, const real *m, const real *v) {   /* row-major 3x3 */
  real x = m[0]*v[0]+m[1]*v[1]+m[2]*v[2], y = m[3]*v[0]+m[4]*v[1]+m[5]*v[2], z = m[6]*v[0]+m[7]*v[1]+m[8]*v[2];
  r[0]=x; r[1]=y; r[2]=z;
}
static inline void matT_vec(real *r, const real *m, const real *v) {
  real x = m[0]*v[0]+m[3]*v[1]+m[6]*v[2], y = m[1]*v[0]+m[4]*v[1]+m[7]*v[2], z = m[2]*v[0]+m[5]*v[1]+m[8]*v[2];
  r[0]=x; r[1]=y; r[2]=z;
}
static void mat_mul(real *r, const real *a, const real *b) {
  real t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    t[3*i+j] = a[3*i]*b[j] + a[3*i+1]*b[3+j] + a[3*i+2]*b[6+j];
  memcpy(r, t, sizeof t);
}
static void quat_mul(real *r, const real *a, const real *b) {
  real w = a[0]*b[0]-a[1]*b[1]-a[2]*b[2]-a[3]*b[3];
  real x = a[0]*b[1]+a[1]*b[0]+a[2]*b[3]-a[3]*b[2];
  real y = a[0]*b[2]-a[1]*b[3]+a[2]*b[0]+a[3]*b[1];
  real z = a[0]*b[3]+a[1]*b[2]-a[2]*b[1]+a[3]*b[0];
  r[0]=w; r[1]=x; r[2]=y; r[3]=z;
}
static void quat_to_mat(real *m, const real *q) {
  real w=q[0], x=q[1], y=q[2], z=q[3];
  m[0]=w*w+x*x-y*y-z*z; m[1]=2*(x*y-w*z);       m[2]=2*(x*z+w*y);
  m[3]=2*(x*y+w*z);     m[4]=w*w-x*x+y*y-z*z;   m[5]=2*(y*z-w*x);
  m[6]=2*(x*z-w*y);     m[7]=2*(y*z+w*x);       m[8]=w*w-x*x-y*y+z*z;
}

/* ------------------------------------------------------------------ B.2 kinematics */
static void kinematics(const omodel *m, odata *d) {
  d->xpos[0][0]=d->xpos[0][1]=d->xpos[0][2]=0;
  d->xquat[0][0]=1; d->xquat[0][1]=d->xquat[0][2]=d->xquat[0][3]=0;
  quat_to_mat(d->xmat[0], d->xquat[0]);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parent[b], j = m->body_jnt[b];
    real bp[3], bq[4];
    for (int k=0;k<3;k++) bp[k]=(real)m->body_pos[b][k];
    for (int k=0;k<4;k++) bq[k]=(real)m->body_quat[b][k];
    if (j >= 0 && m->jnt_type[j] == J_FREE) {
      int a = m->jnt_qposadr[j];
      /* mj_kinematics normalises the free-joint quaternion in qpos */
      real n = RSQRT(d->qpos[a+3]*d->qpos[a+3]+d->qpos[a+4]*d->qpos[a+4]+d->qpos[a+5]*d->qpos[a+5]+d->qpos[a+6]*d->qpos[a+6]);
      for (int k=0;k<4;k++) d->qpos[a+3+k] /= n;
      copy3(d->xpos[b], d->qpos + a);
      for (int k=0;k<4;k++) d->xquat[b][k] = d->qpos[a+3+k];
      copy3(d->xanchor[j], d->xpos[b]);
      d->xaxis[j][0]=0; d->xaxis[j][1]=0; d->xaxis[j][2]=1;
    } else {
      real t[3];
      mat_vec(t, d->xmat[p], bp);
      add3(d->xpos[b], d->xpos[p], t);
      quat_mul(d->xquat[b], d->xquat[p], bq);
      if (j >= 0) {   /* hinge at the body origin (jnt_pos = 0 in the supported scenes) */
        real mt[9], ax[3], qj[4];
        for (int k=0;k<3;k++) ax[k]=(real)m->jnt_axis[j][k];
        quat_to_mat(mt, d->xquat[b]);
        mat_vec(d->xaxis[j], mt, ax);
        copy3(d->xanchor[j], d->xpos[b]);
        real ang = d->qpos[m->jnt_qposadr[j]];
        real s = RSIN(ang*(real)0.5);
        qj[0]=RCOS(ang*(real)0.5); qj[1]=s*ax[0]; qj[2]=s*ax[1]; qj[3]=s*ax[2];
        quat_mul(d->xquat[b], d->xquat[b], qj);
      }
    }
    quat_to_mat(d->xmat[b], d->xquat[b]);
    real ip[3], t[3];
    for (int k=0;k<3;k++) ip[k]=(real)m->body_ipos[b][k];
    mat_vec(t, d->xmat[b], ip);
    add3(d->xipos[b], d->xpos[b], t);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_body[g];
    real gp[3], gq[4], gm[9], t[3];
    for (int k=0;k<3;k++) gp[k]=(real)m->geom_pos[g][k];
    for (int k=0;k<4;k++) gq[k]=(real)m->geom_quat[g][k];
    mat_vec(t, d->xmat[b], gp);
    add3(d->gpos[g], d->xpos[b], t);
    quat_to_mat(gm, gq);
    mat_mul(d->gmat[g], d->xmat[b], gm);
  }
  {
    real sp[3], t[3];
    for (int k=0;k<3;k++) sp[k]=(real)m->tcp_pos[k];
    mat_vec(t, d->xmat[m->tcp_body], sp);
    add3(d->site_tcp, d->xpos[m->tcp_body], t);
  }
}

/* ------------------------------------------------------------------ B.2 com_pos: subtree_com, cinert (BD.2), cdof (BD.1) */
static void com_pos(const omodel *m, odata *d) {
  real mass[MAXB];
  for (int b = 0; b < m->nbody; b++) {
    mass[b] = (real)m->body_mass[b];
    scl3(d->subtree_com[b], d->xipos[b], mass[b]);
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parent[b];
    add3(d->subtree_com[p], d->subtree_com[p], d->subtree_com[b]);
    mass[p] += mass[b];
  }
  for (int b = 0; b < m->nbody; b++) {
    if (mass[b] > 0) scl3(d->subtree_com[b], d->subtree_com[b], 1/mass[b]);
    else copy3(d->subtree_com[b], d->xipos[b]);
  }
  for (int b = 1; b < m->nbody; b++) {
    const real *c = d->subtree_com[m->body_rootid[b]];
    real I[9], R[9], T[9], dd[3], ms = (real)m->body_mass[b];
    for (int k=0;k<9;k++) I[k]=(real)m->body_inertia[b][k];
    memcpy(R, d->xmat[b], sizeof R);
    mat_mul(T, R, I);
    /* T = R I R^T */
    real Rt[9] = {R[0],R[3],R[6],R[1],R[4],R[7],R[2],R[5],R[8]};
    mat_mul(T, T, Rt);
    sub3(dd, d->xipos[b], c);
    real *ci = d->cinert[b];
    ci[0] = T[0] + ms*(dd[1]*dd[1]+dd[2]*dd[2]);
    ci[1] = T[4] + ms*(dd[0]*dd[0]+dd[2]*dd[2]);
    ci[2] = T[8] + ms*(dd[0]*dd[0]+dd[1]*dd[1]);
    ci[3] = T[1] - ms*dd[0]*dd[1];
    ci[4] = T[2] - ms*dd[0]*dd[2];
    ci[5] = T[5] - ms*dd[1]*dd[2];
    ci[6] = ms*dd[0]; ci[7] = ms*dd[1]; ci[8] = ms*dd[2]; ci[9] = ms;
  }
  memset(d->cinert[0], 0, sizeof d->cinert[0]);
  /* dof bookkeeping + cdof */
  int last_dof[MAXB];
  last_dof[0] = -1;
  for (int b = 1; b < m->nbody; b++) {
    int j = m->body_jnt[b];
    last_dof[b] = last_dof[m->body_parent[b]];
    if (j < 0) continue;
    const real *c = d->subtree_com[m->body_rootid[b]];
    int a = m->jnt_dofadr[j];
    if (m->jnt_type[j] == J_HINGE) {
      real off[3];
      sub3(off, c, d->xanchor[j]);
      copy3(d->cdof[a], d->xaxis[j]);
      cross3(d->cdof[a] + 3, d->xaxis[j], off);
      d->dof_body[a] = b; d->dof_parent[a] = last_dof[b]; last_dof[b] = a;
    } else {
      real off[3];
      sub3(off, c, d->xpos[b]);
      for (int k = 0; k < 3; k++) {
        real *cd = d->cdof[a+k];
        cd[0]=cd[1]=cd[2]=0; cd[3]=(k==0); cd[4]=(k==1); cd[5]=(k==2);
        real r[3] = {d->xmat[b][k], d->xmat[b][3+k], d->xmat[b][6+k]};   /* column k */
        real *ca = d->cdof[a+3+k];
        copy3(ca, r);
        cross3(ca + 3, r, off);
      }
      for (int k = 0; k < 6; k++) {
        d->dof_body[a+k] = b; d->dof_parent[a+k] = last_dof[b]; last_dof[b] = a+k;
      }
    }
  }
}

static void mul_inert_vec(real *r, const real *i, const real *v) {   /* BD.2 */
  r[0]=i[0]*v[0]+i[3]*v[1]+i[4]*v[2]-i[8]*v[4]+i[7]*v[5];
  r[1]=i[3]*v[0]+i[1]*v[1]+i[5]*v[2]+i[8]*v[3]-i[6]*v[5];
  r[2]=i[4]*v[0]+i[5]*v[1]+i[2]*v[2]-i[7]*v[3]+i[6]*v[4];
  r[3]=i[8]*v[1]-i[7]*v[2]+i[9]*v[3];
  r[4]=i[6]*v[2]-i[8]*v[0]+i[9]*v[4];
  r[5]=i[7]*v[0]-i[6]*v[1]+i[9]*v[5];
}
static real dot6(const real *a, const real *b) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]+a[3]*b[3]+a[4]*b[4]+a[5]*b[5]; }

/* ------------------------------------------------------------------ BD.3 CRBA + dense Cholesky */
static void crb(const omodel *m, odata *d) {
  int nv = m->nv;
  memcpy(d->crb, d->cinert, sizeof d->crb);
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parent[b];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[p][k] += d->crb[b][k];
  }
  for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) d->M[i][j] = 0;
  for (int i = 0; i < nv; i++) {
    real buf[6];
    mul_inert_vec(buf, d->crb[d->dof_body[i]], d->cdof[i]);
    int jn = -1;
    for (int j = 0; j < m->njnt; j++) {
      int w = m->jnt_type[j] == J_FREE ? 6 : 1;
      if (i >= m->jnt_dofadr[j] && i < m->jnt_dofadr[j] + w) jn = j;
    }
    d->M[i][i] = (real)m->jnt_armature[jn] + dot6(d->cdof[i], buf);
    for (int j = d->dof_parent[i]; j >= 0; j = d->dof_parent[j]) {
      real v = dot6(d->cdof[j], buf);
      d->M[i][j] = v; d->M[j][i] = v;
    }
  }
}
/* lower Cholesky A = L L^T of an n x n SPD matrix stored with stride MAXV */
static void chol(int n, real A[MAXV][MAXV], real L[MAXV][MAXV]) {
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
    real s = A[i][j];
    for (int k = 0; k < j; k++) s -= L[i][k]*L[j][k];
    L[i][j] = (i == j) ? RSQRT(s) : s / L[j][j];
  }
}
static void chol_solve(int n, real L[MAXV][MAXV], real *x, const real *b) {
  real y[MAXV];
  for (int i = 0; i < n; i++) { real s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k]*y[k]; y[i] = s / L[i][i]; }
  for (int i = n-1; i >= 0; i--) { real s = y[i]; for (int k = i+1; k < n; k++) s -= L[k][i]*x[k]; x[i] = s / L[i][i]; }
}
static void mul_M(int n, real M[MAXV][MAXV], real *r, const real *v) {
  for (int i = 0; i < n; i++) { real s = 0; for (int j = 0; j < n; j++) s += M[i][j]*v[j]; r[i] = s; }
}

/* ------------------------------------------------------------------ BD.6 point Jacobian */
static void jac_point(const omodel *m, const odata *d, const real *p, int body, real jacp[MAXV][3]) {
  for (int i = 0; i < m->nv; i++) jacp[i][0]=jacp[i][1]=jacp[i][2]=0;
  if (body <= 0) return;
  real off[3];
  sub3(off, p, d->subtree_com[m->body_rootid[body]]);
  /* dofs of `body` and all its ancestors */
  int b = body;
  while (b > 0) {
    int j = m->body_jnt[b];
    if (j >= 0) {
      int w = m->jnt_type[j] == J_FREE ? 6 : 1, a = m->jnt_dofadr[j];
      for (int k = 0; k < w; k++) {
        real t[3];
        cross3(t, d->cdof[a+k], off);
        add3(jacp[a+k], d->cdof[a+k] + 3, t);
      }
    }
    b = m->body_parent[b];
  }
}

/* ------------------------------------------------------------------ BD.4 com_vel, B.5 passive, BD.5 rne */
static void cross_motion(real *r, const real *v, const real *s) {
  real a[3], b[3], c[3];
  cross3(a, v, s); cross3(b, v, s + 3); cross3(c, v + 3, s);
  r[0]=a[0]; r[1]=a[1]; r[2]=a[2]; r[3]=b[0]+c[0]; r[4]=b[1]+c[1]; r[5]=b[2]+c[2];
}
static void cross_force(real *r, const real *v, const real *f) {
  real a[3], b[3], c[3];
  cross3(a, v, f); cross3(b, v + 3, f + 3); cross3(c, v, f + 3);
  r[0]=a[0]+b[0]; r[1]=a[1]+b[1]; r[2]=a[2]+b[2]; r[3]=c[0]; r[4]=c[1]; r[5]=c[2];
}
static void com_vel(const omodel *m, odata *d) {
  memset(d->cvel[0], 0, sizeof d->cvel[0]);
  for (int b = 1; b < m->nbody; b++) {
    real v[6];
    memcpy(v, d->cvel[m->body_parent[b]], sizeof v);
    int j = m->body_jnt[b];
    if (j >= 0) {
      int a = m->jnt_dofadr[j];
      if (m->jnt_type[j] == J_HINGE) {
        cross_motion(d->cdof_dot[a], v, d->cdof[a]);
        for (int k = 0; k < 6; k++) v[k] += d->cdof[a][k] * d->qvel[a];
      } else {
        for (int i = 0; i < 3; i++) { memset(d->cdof_dot[a+i], 0, sizeof d->cdof_dot[0]);
          for (int k = 0; k < 6; k++) v[k] += d->cdof[a+i][k] * d->qvel[a+i]; }
        for (int i = 3; i < 6; i++) cross_motion(d->cdof_dot[a+i], v, d->cdof[a+i]);
        for (int i = 3; i < 6; i++) for (int k = 0; k < 6; k++) v[k] += d->cdof[a+i][k] * d->qvel[a+i];
      }
    }
    memcpy(d->cvel[b], v, sizeof v);
  }
}
static void passive(const omodel *m, odata *d) {
  for (int i = 0; i < m->nv; i++) d->qfrc_passive[i] = 0;
  for (int j = 0; j < m->njnt; j++) {
    int w = m->jnt_type[j] == J_FREE ? 6 : 1, a = m->jnt_dofadr[j];
    for (int k = 0; k < w; k++) d->qfrc_passive[a+k] -= (real)m->jnt_damping[j] * d->qvel[a+k];
  }
  for (int b = 1; b < m->nbody; b++) {
    if (m->body_gravcomp[b] == 0) continue;
    real jacp[MAXV][3], f[3];
    for (int k = 0; k < 3; k++) f[k] = -(real)m->gravity[k] * (real)m->body_mass[b] * (real)m->body_gravcomp[b];
    jac_point(m, d, d->xipos[b], b, jacp);
    for (int i = 0; i < m->nv; i++) d->qfrc_passive[i] += dot3(jacp[i], f);
  }
}
static void rne(const omodel *m, odata *d) {
  real cacc[MAXB][6], cfrc[MAXB][6];
  cacc[0][0]=cacc[0][1]=cacc[0][2]=0;
  for (int k = 0; k < 3; k++) cacc[0][3+k] = -(real)m->gravity[k];
  memset(cfrc[0], 0, sizeof cfrc[0]);
  for (int b = 1; b < m->nbody; b++) {
    memcpy(cacc[b], cacc[m->body_parent[b]], sizeof cacc[0]);
    int j = m->body_jnt[b];
    if (j >= 0) {
      int w = m->jnt_type[j] == J_FREE ? 6 : 1, a = m->jnt_dofadr[j];
      for (int i = 0; i < w; i++) for (int k = 0; k < 6; k++) cacc[b][k] += d->cdof_dot[a+i][k] * d->qvel[a+i];
    }
    real t[6], t2[6];
    mul_inert_vec(cfrc[b], d->cinert[b], cacc[b]);
    mul_inert_vec(t, d->cinert[b], d->cvel[b]);
    cross_force(t2, d->cvel[b], t);
    for (int k = 0; k < 6; k++) cfrc[b][k] += t2[k];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parent[b];
    for (int k = 0; k < 6; k++) cfrc[p][k] += cfrc[b][k];
  }
  for (int i = 0; i < m->nv; i++) d->qfrc_bias[i] = dot6(d->cdof[i], cfrc[d->dof_body[i]]);
}

/* ------------------------------------------------------------------ B.3 / BD.12 colliders */
/* MJX math.orthogonals + make_frame: rows n, t1, t2 */
static void make_frame(real *frame, const real *nin) {
  real n[3] = {nin[0], nin[1], nin[2]};
  normalize3(n);
  real b[3] = {0, 0, 0};
  if (n[1] > (real)-0.5 && n[1] < (real)0.5) b[1] = 1; else b[2] = 1;
  real s = dot3(n, b);
  addscl3(b, b, n, -s);
  normalize3(b);
  if (n[0] == 0 && n[1] == 0 && n[2] == 0) b[0]=b[1]=b[2]=0;
  copy3(frame, n); copy3(frame + 3, b); cross3(frame + 6, n, b);
}

/* MJX collision_primitive.plane_capsule */
static void plane_capsule(const real *ppos, const real *pmat, const real *cpos, const real *cmat, const real *csize,
                          real *dist, real pos[][3], real frame[][9]) {
  real n[3] = {pmat[2], pmat[5], pmat[8]}, axis[3] = {cmat[2], cmat[5], cmat[8]};
  real b[3], fr[9], e[2][3];
  for (int k = 0; k < 2; k++) {
    real t[3], sgn = k == 0 ? 1 : -1;
    addscl3(e[k], cpos, axis, sgn * csize[1]);
    sub3(t, e[k], ppos);
    dist[k] = dot3(t, n) - csize[0];
  }
  OINACTIVE_BEGIN(dist[0] < 0 || dist[1] < 0);
  addscl3(b, axis, n, -dot3(n, axis));
  real bn = normalize3(b);
  if (bn < (real)0.5) { b[0]=0; if (n[1] > (real)-0.5 && n[1] < (real)0.5) { b[1]=1; b[2]=0; } else { b[1]=0; b[2]=1; } }
  copy3(fr, n); copy3(fr + 3, b); cross3(fr + 6, n, b);
  for (int k = 0; k < 2; k++) {
    addscl3(pos[k], e[k], n, -(csize[0] + (real)0.5 * dist[k]));
    memcpy(frame[k], fr, sizeof fr);
  }
  OINACTIVE_END();
}

/* MJX math.closest_segment_point_and_dist */
static real closest_segment_point(real *res, const real *a, const real *b, const real *pt) {
  real ab[3], ap[3], t, d[3];
  sub3(ab, b, a); sub3(ap, pt, a);
  t = dot3(ap, ab) / (dot3(ab, ab) + (real)1e-6);
  t = t < 0 ? (real)0 : (t > 1 ? (real)1 : t);
  addscl3(res, a, ab, t);
  sub3(d, pt, res);
  return dot3(d, d);
}
/* MJX math.closest_segment_to_segment_points */
static void closest_seg_seg(real *besta, real *bestb, const real *a0, const real *a1, const real *b0, const real *b1) {
  real da[3], db[3], amid[3], bmid[3], tr[3];
  sub3(da, a1, a0); sub3(db, b1, b0);
  real la = normalize3(da), lb = normalize3(db);
  real ha = la * (real)0.5, hb = lb * (real)0.5;
  addscl3(amid, a0, da, ha); addscl3(bmid, b0, db, hb);
  sub3(tr, amid, bmid);
  real dd = dot3(da, db), dat = dot3(da, tr), dbt = dot3(db, tr);
  real denom = 1 - dd * dd;
  real ta = (-dat + dd * dbt) / (denom + (real)1e-6);
  real tb = dbt + ta * dd;
  ta = ta < -ha ? -ha : (ta > ha ? ha : ta);
  tb = tb < -hb ? -hb : (tb > hb ? hb : tb);
  addscl3(besta, amid, da, ta); addscl3(bestb, bmid, db, tb);
  real na[3], nb[3];
  real d1 = closest_segment_point(na, a0, a1, bestb);
  real d2 = closest_segment_point(nb, b0, b1, besta);
  if (d1 < d2) copy3(besta, na); else copy3(bestb, nb);
}
/* MJX collision_primitive.capsule_capsule -> _sphere_sphere */
static void capsule_capsule(const real *p1, const real *m1, const real *s1, const real *p2, const real *m2, const real *s2,
                            real *dist, real pos[][3], real frame[][9]) {
  real ax1[3] = {m1[2], m1[5], m1[8]}, ax2[3] = {m2[2], m2[5], m2[8]};
  real a0[3], a1[3], b0[3], b1[3], pa[3], pb[3], n[3];
  addscl3(a0, p1, ax1, -s1[1]); addscl3(a1, p1, ax1, s1[1]);
  addscl3(b0, p2, ax2, -s2[1]); addscl3(b1, p2, ax2, s2[1]);
  closest_seg_seg(pa, pb, a0, a1, b0, b1);
  sub3(n, pb, pa);
  real dn = normalize3(n);
  if (dn == 0) { n[0]=1; n[1]=0; n[2]=0; }
  dist[0] = dn - (s1[0] + s2[0]);
  OINACTIVE_BEGIN(dist[0] < 0);
  addscl3(pos[0], pa, n, s1[0] + dist[0] * (real)0.5);
  make_frame(frame[0], n);
  OINACTIVE_END();
}

/* box faces in the box frame: index f = 2*axis + (0: +, 1: -).  Vertices counter-clockwise seen
 * from outside so that cross(edge, normal) points out of the face region. */
static void box_face(const real *s, int f, real v[4][3], real *n) {
  int k = f >> 1, u = (k + 1) % 3, w = (k + 2) % 3;
  real sg = (f & 1) ? -1 : 1;
  n[0]=n[1]=n[2]=0; n[k] = sg;
  /* for +k: (u,w) = (-,-),(+,-),(+,+),(-,+) is CCW about +k since u x w = k; reversed for -k */
  static const int su[4] = {-1, 1, 1, -1}, sw[4] = {-1, -1, 1, 1};
  for (int i = 0; i < 4; i++) {
    int ii = (f & 1) ? 3 - i : i;
    v[i][k] = sg * s[k]; v[i][u] = su[ii] * s[u]; v[i][w] = sw[ii] * s[w];
  }
}
/* MJX collision_convex._closest_segment_point_plane */
static void seg_point_plane(real *res, const real *a, const real *b, const real *p0, const real *n) {
  real ab[3];
  sub3(ab, b, a);
  real dd = dot3(p0, n), denom = dot3(n, ab);
  real t = (dd - dot3(n, a)) / (denom + (denom == 0 ? (real)1e-6 : (real)0));
  t = t < 0 ? (real)0 : (t > 1 ? (real)1 : t);
  addscl3(res, a, ab, t);
}
/* MJX collision_convex._clip_edge_to_planes; returns mask */
static int clip_edge_to_planes(const real *p0, const real *p1, int np, real pp[][3], real pn[][3], real out[2][3]) {
  int f0[8], f1[8], any_both = 0;
  real cand[8][3];
  for (int i = 0; i < np; i++) {
    real t[3];
    sub3(t, p0, pp[i]); f0[i] = dot3(t, pn[i]) > (real)1e-6;
    sub3(t, p1, pp[i]); f1[i] = dot3(t, pn[i]) > (real)1e-6;
    seg_point_plane(cand[i], p0, p1, pp[i], pn[i]);
    any_both |= (f0[i] && f1[i]);
  }
  real e01[3], e10[3];
  sub3(e01, p1, p0); sub3(e10, p0, p1);
  real best = 0; int bi = -1;
  for (int i = 0; i < np; i++) {
    real t[3]; const real *e = f0[i] ? cand[i] : p0;
    sub3(t, e, p0);
    real dd = dot3(t, e01);
    if (bi < 0 || dd > best) { best = dd; bi = i; }
  }
  copy3(out[0], f0[bi] ? cand[bi] : p0);
  best = 0; bi = -1;
  for (int i = 0; i < np; i++) {
    real t[3]; const real *e = f1[i] ? cand[i] : p1;
    sub3(t, e, p1);
    real dd = dot3(t, e10);
    if (bi < 0 || dd > best) { best = dd; bi = i; }
  }
  copy3(out[1], f1[bi] ? cand[bi] : p1);
  int mask = !any_both;
  if (!mask) { copy3(out[0], p0); copy3(out[1], p1); }
  real dn[3];
  sub3(dn, out[0], out[1]);
  if (dot3(e10, dn) < 0) mask = 0;
  return mask;
}
/* MJX collision_convex._capsule_convex specialised to a box.  Two restatements
 * selected by omodel.capbox_mode:
 *
 * mode 1 (default) -- mujoco-mjx 3.x as recalled by the round-1 review (VERDICT.md "What's weak" #1) and
 * pinned, weakly, by the reference's own recording (tests/test_oracle_pins.py::test_recorded_run_pins_
 * capsule_box_far_field: replaying data/theta.csv must not produce the +1 -> 0.2 sentinel flips that
 * data/cost_c.csv, max 0.021, never shows):
 *  - support_f = min over the two end points of dot(pt - r n_f - face_f[0], n_f) for each of the 6 faces;
 *    has_support = all(support_f < 0); best face = first argmax of support_f;
 *  - the segment is clipped against the 4 side planes of the best face (_clip_edge_to_planes); the two
 *    clipped points, pushed by the radius along -n, are measured against the face plane;
 *    face_penetration = where(mask & has_support, ., -1)  => both face slots are the dist = +1 sentinel
 *    whenever any face plane separates the radius-inflated capsule from the box: no far-field distance;
 *  - shallow edge contact over all 12 box edges: closest points edge <-> segment, edge_axis = normalised
 *    (edge_pt - cap_pt); an edge qualifies when it is not degenerate (|.|^2 >= 1e-6) and the capsule point
 *    lies in front of both faces adjacent to the edge (edge_voronoi_front); edge_penetration = r - dist for
 *    qualifying edges, -1 otherwise; the edge with the largest penetration (first) is the candidate;
 *    has_edge_contact = pen > 0 & (min face pen > 0 ? pen < min face pen : true) & !(|edge_axis . n| > 0.99);
 *    it replaces slot 0 (position, normal = edge_axis, penetration).
 *
 * mode 0 -- the round-1 restatement (brax-era collider): best face by the same argmax, face slots report
 * the true face distance whenever the clip succeeds (no has_support gate), edge contact = closest of the
 * best face's 4 edges with r - dist > 0.  Kept only so that the pin test can show it contradicts the
 * reference's recording. */
static void capsule_box(const real *cpos, const real *cmat, const real *csize, const real *bpos, const real *bmat, const real *bsize,
                        real *dist, real pos[][3], real frame[][9], int mode) {
  real t[3], cp[3], ax[3], axw[3] = {cmat[2], cmat[5], cmat[8]}, seg[3], pts[2][3];
  const real r = csize[0];
  sub3(t, cpos, bpos); matT_vec(cp, bmat, t);
  matT_vec(ax, bmat, axw);
  scl3(seg, ax, csize[1]);
  sub3(pts[0], cp, seg); add3(pts[1], cp, seg);
  int best = 0, has_support = 1; real bests = 0;
  for (int f = 0; f < 6; f++) {
    int k = f >> 1; real sg = (f & 1) ? -1 : 1;
    real s0 = sg * pts[0][k] - r - bsize[k], s1 = sg * pts[1][k] - r - bsize[k];
    real sup = s0 < s1 ? s0 : s1;
    if (f == 0 || sup > bests) { bests = sup; best = f; }
    if (!(sup < 0)) has_support = 0;
  }
  if (mode == 0) has_support = 1;
  real face[4][3], n[3], ep0[4][3], en[4][3];
  box_face(bsize, best, face, n);
  for (int i = 0; i < 4; i++) {
    copy3(ep0[i], face[(i + 3) % 4]);          /* edge_p0 = roll(face, 1), edge_p1 = face */
    real e[3]; sub3(e, face[i], ep0[i]);
    cross3(en[i], e, n);
  }
  real cl[2][3];
  int mask = clip_edge_to_planes(pts[0], pts[1], 4, ep0, en, cl);
  real nrm[2][3], pen[2], lp[2][3];
  for (int k = 0; k < 2; k++) {
    real c[3], fp[3], tt[3];
    addscl3(c, cl[k], n, -r);
    sub3(tt, c, face[0]);
    addscl3(fp, c, n, -dot3(tt, n));
    for (int q = 0; q < 3; q++) lp[k][q] = (c[q] + fp[q]) * (real)0.5;
    sub3(tt, fp, c);
    pen[k] = (mask && has_support) ? dot3(tt, n) : (real)-1;
    scl3(nrm[k], n, -1);
  }
  if (mode == 0) {
    /* edge contact, round-1 restatement */
    real bd = 0; int be = -1; real bec[3], bcc[3];
    for (int i = 0; i < 4; i++) {
      real ec[3], cc[3], df[3];
      closest_seg_seg(ec, cc, ep0[i], face[i], pts[0], pts[1]);
      sub3(df, ec, cc);
      real dd = dot3(df, df);
      if (be < 0 || dd < bd) { bd = dd; be = i; copy3(bec, ec); copy3(bcc, cc); }
    }
    real eax[3];
    sub3(eax, bcc, bec);
    real ed = normalize3(eax);
    real epen = r - ed;
    if (epen > 0) {
      for (int q = 0; q < 3; q++) lp[0][q] = (bec[q] + (bcc[q] - eax[q] * r)) * (real)0.5;
      scl3(nrm[0], eax, -1);
      pen[0] = epen;
    }
  } else {
    /* shallow edge contact over the 12 edges: edge e = 4*k + 2*iu + iw runs along axis k at
     * (u, w) = (su * s_u, sw * s_w), su = iu ? +1 : -1, sw = iw ? +1 : -1; adjacent faces: su*e_u, sw*e_w */
    real bpen = -1; int be = -1; real beax[3] = {0, 0, 0}, bec[3] = {0, 0, 0}, bcc[3] = {0, 0, 0};
    for (int e = 0; e < 12; e++) {
      int k = e >> 2, u = (k + 1) % 3, w = (k + 2) % 3;
      real su = (e & 2) ? 1 : -1, sw = (e & 1) ? 1 : -1;
      real e0[3], e1[3], ec[3], cc[3], dir[3];
      e0[k] = -bsize[k]; e1[k] = bsize[k]; e0[u] = e1[u] = su * bsize[u]; e0[w] = e1[w] = sw * bsize[w];
      closest_seg_seg(ec, cc, e0, e1, pts[0], pts[1]);
      sub3(dir, ec, cc);
      int degenerate = dot3(dir, dir) < (real)1e-6;
      real ed = normalize3(dir);
      int front = (su * dir[u] < 0) && (sw * dir[w] < 0);
      real epen = (!degenerate && front) ? r - ed : (real)-1;
      if (be < 0 || epen > bpen) { bpen = epen; be = e; copy3(beax, dir); copy3(bec, ec); copy3(bcc, cc); }
    }
    int degenerate = 0;
    { real d[3]; sub3(d, bec, bcc); degenerate = dot3(d, d) < (real)1e-6; }
    int parallel = (RFABS(dot3(beax, n)) > (real)0.99) && !degenerate;
    real minface = pen[0] < pen[1] ? pen[0] : pen[1];
    int has_edge = (bpen > 0) && (minface > 0 ? bpen < minface : 1) && !parallel;
    if (has_edge) {
      for (int q = 0; q < 3; q++) lp[0][q] = (bec[q] + (bcc[q] + beax[q] * r)) * (real)0.5;
      copy3(nrm[0], beax);
      pen[0] = bpen;
    }
  }
  for (int k = 0; k < 2; k++) {
    real w[3], nw[3];
    mat_vec(w, bmat, lp[k]); add3(pos[k], w, bpos);
    mat_vec(nw, bmat, nrm[k]);
    make_frame(frame[k], nw);
    dist[k] = -pen[k];
  }
}

/* MJX collision_convex._manifold_points: 4 points of (approximately) maximal area */
static void manifold_points(int n, real poly[][3], const int *mask, const real *nrm, int idx[4]) {
  real dm[16];
  for (int i = 0; i < n; i++) dm[i] = mask[i] ? (real)0 : (real)-1e6;
  int a = 0;
  for (int i = 1; i < n; i++) if (dm[i] > dm[a]) a = i;
  int b = 0; real bv = 0;
  for (int i = 0; i < n; i++) { real t[3]; sub3(t, poly[a], poly[i]); real v = dot3(t, t) + dm[i]; if (i == 0 || v > bv) { bv = v; b = i; } }
  real ab[3], t[3];
  sub3(t, poly[a], poly[b]); cross3(ab, nrm, t);
  int c = 0; real cv = 0;
  for (int i = 0; i < n; i++) { real ap[3]; sub3(ap, poly[a], poly[i]); real v = RFABS(dot3(ap, ab)) + dm[i]; if (i == 0 || v > cv) { cv = v; c = i; } }
  real ac[3], bc[3];
  sub3(t, poly[a], poly[c]); cross3(ac, nrm, t);
  sub3(t, poly[b], poly[c]); cross3(bc, nrm, t);
  /* d: furthest from the two other triangle edges.  Deviation from the MJX tie rule (first
   * argmax, which on an exact square re-selects `a`): already chosen points are penalised so a
   * fourth distinct corner is returned whenever one exists. */
  int dsel = 0; real dv = 0;
  for (int i = 0; i < n; i++) {
    real bp[3], ap[3];
    sub3(bp, poly[b], poly[i]); sub3(ap, poly[a], poly[i]);
    real v1 = RFABS(dot3(bp, bc)), v2 = RFABS(dot3(ap, ac));
    real v = (v1 > v2 ? v1 : v2) + dm[i] - ((i == a || i == b || i == c) ? (real)2e6 : (real)0);
    if (i == 0 || v > dv) { dv = v; dsel = i; }
  }
  idx[0]=a; idx[1]=b; idx[2]=c; idx[3]=dsel;
}
static void box_verts(const real *s, real v[8][3]) {
  for (int i = 0; i < 8; i++) { v[i][0] = (i & 4 ? 1 : -1) * s[0]; v[i][1] = (i & 2 ? 1 : -1) * s[1]; v[i][2] = (i & 1 ? 1 : -1) * s[2]; }
}
/* MJX collision_convex.plane_convex for a box */
static void plane_box(const real *ppos, const real *pmat, const real *bpos, const real *bmat, const real *bsize,
                      real *dist, real pos[][3], real frame[][9]) {
  real v[8][3], t[3], pl[3], n[3], nw[3] = {pmat[2], pmat[5], pmat[8]}, sup[8], smax = 0;
  int mask[8], idx[4];
  box_verts(bsize, v);
  sub3(t, ppos, bpos); matT_vec(pl, bmat, t); matT_vec(n, bmat, nw);
  for (int i = 0; i < 8; i++) { sub3(t, pl, v[i]); sup[i] = dot3(t, n); if (i == 0 || sup[i] > smax) smax = sup[i]; }
  real thr = smax - (real)1e-3; if (thr < 0) thr = 0;
  for (int i = 0; i < 8; i++) mask[i] = sup[i] > thr;
  manifold_points(8, v, mask, n, idx);
  real fr[9];
  make_frame(fr, nw);
  for (int k = 0; k < 4; k++) {
    int uniq = 1;
    for (int q = 0; q < k; q++) if (idx[q] == idx[k]) uniq = 0;
    real w[3];
    mat_vec(w, bmat, v[idx[k]]); add3(w, w, bpos);
    dist[k] = uniq ? -sup[idx[k]] : (real)1;
    addscl3(pos[k], w, nw, -(real)0.5 * dist[k]);
    memcpy(frame[k], fr, sizeof fr);
  }
}

/* Box-box: separating-axis test over the 15 axes, then either a clipped face-face manifold reduced
 * to 4 points (MJX _create_contact_manifold + _manifold_points) or a single edge-edge contact.
 * Only target_0 vs the static boxes uses this; those slots are not robot-involving and never enter
 * the cost, so only penetrating configurations matter. */
static int clip_poly_halfplane(int n, real in[][3], real out[][3], const real *pp, const real *pn) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    const real *a = in[i], *b = in[(i + 1) % n];
    real ta[3], tb[3];
    sub3(ta, a, pp); sub3(tb, b, pp);
    real da = dot3(ta, pn), db = dot3(tb, pn);
    if (da <= 0) { copy3(out[m], a); m++; }
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      real s = da / (da - db), ab[3];
      sub3(ab, b, a); addscl3(out[m], a, ab, s); m++;
    }
  }
  return m;
}
static void box_box(const real *p1, const real *m1, const real *s1, const real *p2, const real *m2, const real *s2,
                    real *dist, real pos[][3], real frame[][9]) {
  /* work in the frame of box 2 */
  real t[3], c[3], R[9];
  sub3(t, p1, p2); matT_vec(c, m2, t);
  { real m2t[9] = {m2[0],m2[3],m2[6],m2[1],m2[4],m2[7],m2[2],m2[5],m2[8]}; mat_mul(R, m2t, m1); }
  /* axes of box1 in frame 2 = columns of R */
  real A[3][3];
  for (int i = 0; i < 3; i++) { A[i][0]=R[i]; A[i][1]=R[3+i]; A[i][2]=R[6+i]; }
  real bestsep = -1e30; int besttype = -1, bi = 0, bj = 0; real bestn[3] = {0,0,1};
  /* face axes of box 2 (type 0), box 1 (type 1), edge x edge (type 2); normal oriented from box1 to box2 */
  for (int type = 0; type < 3; type++) for (int i = 0; i < 3; i++) for (int j = 0; j < (type == 2 ? 3 : 1); j++) {
    real ax[3] = {0,0,0};
    if (type == 0) ax[i] = 1; else if (type == 1) copy3(ax, A[i]);
    else { real e2[3] = {0,0,0}; e2[j] = 1; cross3(ax, A[i], e2); if (normalize3(ax) < (real)1e-6) continue; }
    real r1 = 0, r2 = 0;
    for (int k = 0; k < 3; k++) { r1 += s1[k] * RFABS(dot3(A[k], ax)); r2 += s2[k] * RFABS(ax[k]); }
    real dc = dot3(c, ax);
    real sep = RFABS(dc) - r1 - r2;
    /* prefer face axes over edge axes on near ties (standard bias) */
    real cmp = sep - (type == 2 ? (real)1e-6 : (real)0);
    if (cmp > bestsep) {
      bestsep = cmp; besttype = type; bi = i; bj = j;
      real sg = dc > 0 ? -1 : 1;           /* from box1 (at c) towards box2 (at origin) */
      scl3(bestn, ax, sg);
    }
  }
  for (int k = 0; k < 4; k++) { dist[k] = 1; pos[k][0]=pos[k][1]=pos[k][2]=0; }
  real nw[3];
  mat_vec(nw, m2, bestn);
  for (int k = 0; k < 4; k++) make_frame(frame[k], nw);
  if (besttype == 2) {
    /* edge-edge: support edges along bestn */
    real e1c[3], e2c[3] = {0,0,0};
    copy3(e1c, c);
    for (int k = 0; k < 3; k++) if (k != bi) { real sgn = dot3(A[k], bestn) > 0 ? 1 : -1; addscl3(e1c, e1c, A[k], sgn * s1[k]); }
    for (int k = 0; k < 3; k++) if (k != bj) { real sgn = bestn[k] > 0 ? -1 : 1; e2c[k] = sgn * s2[k]; }
    real a0[3], a1[3], b0[3], b1[3], pa[3], pb[3], e2[3] = {0,0,0};
    e2[bj] = 1;
    addscl3(a0, e1c, A[bi], -s1[bi]); addscl3(a1, e1c, A[bi], s1[bi]);
    addscl3(b0, e2c, e2, -s2[bj]); addscl3(b1, e2c, e2, s2[bj]);
    closest_seg_seg(pa, pb, a0, a1, b0, b1);
    real mid[3] = {(pa[0]+pb[0])*(real)0.5, (pa[1]+pb[1])*(real)0.5, (pa[2]+pb[2])*(real)0.5}, w[3];
    mat_vec(w, m2, mid); add3(pos[0], w, p2);
    real df[3]; sub3(df, pb, pa);
    dist[0] = dot3(df, bestn);
    return;
  }
  /* face-face: reference face on the box owning the axis, incident face on the other box.
   * Do the clipping in the frame of the reference box. */
  real rc[3], Rr[9], rs[3], is[3], nref[3];
  int swap = besttype == 1;     /* reference = box1 */
  if (!swap) {  /* reference box2 (frame 2): incident box1 at c with rotation R; contact normal n points 1->2, ref outward normal = -n */
    copy3(rc, c); memcpy(Rr, R, sizeof R); copy3(rs, s2); copy3(is, s1); scl3(nref, bestn, -1);
  } else {      /* reference box1: express box2 in frame 1 */
    real Rt[9] = {R[0],R[3],R[6],R[1],R[4],R[7],R[2],R[5],R[8]}, mc[3] = {-c[0], -c[1], -c[2]};
    mat_vec(rc, Rt, mc); memcpy(Rr, Rt, sizeof Rt); copy3(rs, s1); copy3(is, s2);
    mat_vec(nref, Rt, bestn);       /* outward normal of the reference face, in frame 1: +n */
  }
  /* reference face: axis k with the largest |nref| */
  int k = 0; for (int q = 1; q < 3; q++) if (RFABS(nref[q]) > RFABS(nref[k])) k = q;
  int rf = 2 * k + (nref[k] > 0 ? 0 : 1);
  real rface[4][3], rn[3];
  box_face(rs, rf, rface, rn);
  /* incident face: the face of the incident box most anti-parallel to rn */
  int inf = 0; real mind = 1e30;
  for (int f = 0; f < 6; f++) {
    int kk = f >> 1; real sg = (f & 1) ? -1 : 1;
    real fn[3] = {sg * Rr[kk], sg * Rr[3+kk], sg * Rr[6+kk]};
    real dd = dot3(fn, rn);
    if (dd < mind) { mind = dd; inf = f; }
  }
  real iface[4][3], itmp[3];
  box_face(is, inf, iface, itmp);
  real poly[16][3], poly2[16][3];
  for (int i = 0; i < 4; i++) { mat_vec(poly[i], Rr, iface[i]); add3(poly[i], poly[i], rc); }
  int np = 4;
  for (int i = 0; i < 4 && np > 0; i++) {
    real e[3], en[3];
    sub3(e, rface[i], rface[(i + 3) % 4]);
    cross3(en, e, rn); normalize3(en);
    np = clip_poly_halfplane(np, poly, poly2, rface[i], en);
    memcpy(poly, poly2, sizeof(real) * 3 * np);
  }
  if (np == 0) return;
  real pref[16][3]; int mask[16], idx[4];
  real depth[16];
  for (int i = 0; i < np; i++) {
    real tt[3]; sub3(tt, poly[i], rface[0]);
    real h = dot3(tt, rn);            /* height above the reference face; penetrating if < 0 */
    depth[i] = -h; mask[i] = h < 0;
    addscl3(pref[i], poly[i], rn, -h);
  }
  manifold_points(np, pref, mask, rn, idx);
  for (int q = 0; q < 4; q++) {
    int i = idx[q], uniq = 1;
    for (int z = 0; z < q; z++) if (idx[z] == i) uniq = 0;
    if (!mask[i] || !uniq) continue;
    real w[3];
    /* back to frame 2 then world */
    if (swap) { real u[3]; mat_vec(u, R, pref[i]); add3(u, u, c); mat_vec(w, m2, u); }
    else mat_vec(w, m2, pref[i]);
    add3(pos[q], w, p2);
    dist[q] = -depth[i];
  }
}

#ifdef ORACLE_COUNT
/* Exact shortcuts taken only by the op-counting build, so that it tallies the work the results depend on rather than
 * MJX's dense evaluation (tests/test_flop_count.py: outputs identical to the dense build, bit for bit):
 *  - capsule vs box: if some box axis separates the segment's bounding interval from the box by >= r, has_support
 *    fails and no edge can be within r, so both slots are the +1 sentinel (rollout_core.h capbox_far);
 *  - box vs box: disjoint world AABBs cannot have an active slot (these slots never enter the planner's cost). */
static int count_capbox_far(const real *cpos, const real *cmat, const real *csize, const real *bpos, const real *bmat, const real *bsize) {
  real t[3], cp[3], ax[3], axw[3] = {cmat[2], cmat[5], cmat[8]};
  sub3(t, cpos, bpos); matT_vec(cp, bmat, t);
  matT_vec(ax, bmat, axw);
  for (int k = 0; k < 3; k++) {
    real h = RFABS(ax[k] * csize[1]);
    real lo = cp[k] - h, hi = cp[k] + h;
    real sep = (lo > -hi ? lo : -hi) - bsize[k];
    if (!(sep < csize[0])) return 1;
  }
  return 0;
}
static int count_aabb_disjoint(const real *p1, const real *m1, const real *s1, const real *p2, const real *m2, const real *s2) {
  for (int i = 0; i < 3; i++) {
    real e1 = RFABS(m1[3*i]) * s1[0] + RFABS(m1[3*i+1]) * s1[1] + RFABS(m1[3*i+2]) * s1[2];
    real e2 = RFABS(m2[3*i]) * s2[0] + RFABS(m2[3*i+1]) * s2[1] + RFABS(m2[3*i+2]) * s2[2];
    if (RFABS(p1[i] - p2[i]) > e1 + e2) return 1;
  }
  return 0;
}
#endif
static void collision(const omodel *m, odata *d) {
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_g1[p], g2 = m->pair_g2[p], a = m->pair_slotadr[p];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    real s1[3], s2[3];
    for (int k = 0; k < 3; k++) { s1[k] = (real)m->geom_size[g1][k]; s2[k] = (real)m->geom_size[g2][k]; }
#ifdef ORACLE_COUNT
    if ((t1 == G_CAPSULE && t2 == G_BOX && count_capbox_far(d->gpos[g1], d->gmat[g1], s1, d->gpos[g2], d->gmat[g2], s2)) ||
        (t1 == G_BOX && t2 == G_BOX && count_aabb_disjoint(d->gpos[g1], d->gmat[g1], s1, d->gpos[g2], d->gmat[g2], s2))) {
      for (int k = 0; k < m->pair_nslot[p]; k++) d->con_dist[a+k] = 1;
    } else
#endif
    if (t1 == G_PLANE && t2 == G_CAPSULE) plane_capsule(d->gpos[g1], d->gmat[g1], d->gpos[g2], d->gmat[g2], s2, d->con_dist + a, d->con_pos + a, d->con_frame + a);
    else if (t1 == G_PLANE && t2 == G_BOX) plane_box(d->gpos[g1], d->gmat[g1], d->gpos[g2], d->gmat[g2], s2, d->con_dist + a, d->con_pos + a, d->con_frame + a);
    else if (t1 == G_CAPSULE && t2 == G_CAPSULE) capsule_capsule(d->gpos[g1], d->gmat[g1], s1, d->gpos[g2], d->gmat[g2], s2, d->con_dist + a, d->con_pos + a, d->con_frame + a);
    else if (t1 == G_CAPSULE && t2 == G_BOX) capsule_box(d->gpos[g1], d->gmat[g1], s1, d->gpos[g2], d->gmat[g2], s2, d->con_dist + a, d->con_pos + a, d->con_frame + a, m->capbox_mode);
    else if (t1 == G_BOX && t2 == G_BOX) box_box(d->gpos[g1], d->gmat[g1], s1, d->gpos[g2], d->gmat[g2], s2, d->con_dist + a, d->con_pos + a, d->con_frame + a);
    for (int k = 0; k < m->pair_nslot[p]; k++) {
      d->con_b1[a+k] = m->geom_body[g1]; d->con_b2[a+k] = m->geom_body[g2];
      d->con_g1[a+k] = g1; d->con_g2[a+k] = g2;
    }
  }
}

/* ------------------------------------------------------------------ B.4 / BD.8 constraints */
static void kbi(const omodel *m, const double *solref, const double *solimp, real pos, real *k, real *b, real *imp) {
  real tc = (real)solref[0], dr = (real)solref[1];
  if (tc < 2 * (real)m->timestep) tc = 2 * (real)m->timestep;
  real dmin = (real)solimp[0], dmax = (real)solimp[1], width = (real)solimp[2], mid = (real)solimp[3], power = (real)solimp[4];
  if (dmin < MJ_MINIMP) dmin = (real)MJ_MINIMP; if (dmin > MJ_MAXIMP) dmin = (real)MJ_MAXIMP;
  if (dmax < MJ_MINIMP) dmax = (real)MJ_MINIMP; if (dmax > MJ_MAXIMP) dmax = (real)MJ_MAXIMP;
  if (width < MJ_MINVAL) width = (real)MJ_MINVAL;
  if (mid < MJ_MINIMP) mid = (real)MJ_MINIMP; if (mid > MJ_MAXIMP) mid = (real)MJ_MAXIMP;
  if (power < 1) power = 1;
  *k = 1 / (dmax * dmax * tc * tc * dr * dr);
  *b = 2 / (dmax * tc);
  if (solref[0] <= 0) *k = -(real)solref[0] / (dmax * dmax);
  if (solref[1] <= 0) *b = -(real)solref[1] / dmax;
  real x = RFABS(pos) / width, y;
  if (x < mid) y = RPOW(x, power) / RPOW(mid, power - 1);
  else y = 1 - RPOW(1 - x, power) / RPOW(1 - mid, power - 1);
  real im = dmin + y * (dmax - dmin);
  if (im < dmin) im = dmin; if (im > dmax) im = dmax;
  if (x > 1) im = dmax;
  *imp = im;
}
static void add_row(const omodel *m, odata *d, const real *J, real pos, real invweight, const double *solref, const double *solimp) {
  int r = d->nefc++;
  real k, b, imp, vel = 0;
  for (int i = 0; i < m->nv; i++) { d->efc_J[r][i] = J[i]; vel += J[i] * d->qvel[i]; }
  kbi(m, solref, solimp, pos, &k, &b, &imp);
  real R = invweight * (1 - imp) / imp;
  if (R < MJ_MINVAL) R = (real)MJ_MINVAL;
  d->efc_D[r] = 1 / R;
  d->efc_aref[r] = -b * vel - k * imp * pos;
  d->efc_pos[r] = pos;
}
static void make_constraint(const omodel *m, odata *d) {
  d->nefc = 0;
  real J[MAXV];
  for (int j = 0; j < m->njnt; j++) {
    if (m->jnt_type[j] != J_HINGE || !m->jnt_limited[j]) continue;
    real q = d->qpos[m->jnt_qposadr[j]];
    real dlo = q - (real)m->jnt_range[j][0], dhi = (real)m->jnt_range[j][1] - q;
    real pos = (dlo < dhi ? dlo : dhi) - (real)m->jnt_margin[j];
    if (!(pos < 0)) continue;
    for (int i = 0; i < m->nv; i++) J[i] = 0;
    J[m->jnt_dofadr[j]] = dlo < dhi ? 1 : -1;
    add_row(m, d, J, pos, (real)m->dof_invweight0[m->jnt_dofadr[j]], m->jnt_solref[j], m->jnt_solimp[j]);
  }
  for (int c = 0; c < m->ncon; c++) {
    if (!(d->con_dist[c] < 0)) continue;
    int b1 = d->con_b1[c], b2 = d->con_b2[c], g1 = d->con_g1[c], g2 = d->con_g2[c];
    real j1[MAXV][3], j2[MAXV][3], dn[MAXV], dt1[MAXV], dt2[MAXV];
    jac_point(m, d, d->con_pos[c], b1, j1);
    jac_point(m, d, d->con_pos[c], b2, j2);
    const real *fr = d->con_frame[c];
    for (int i = 0; i < m->nv; i++) {
      real df[3]; sub3(df, j2[i], j1[i]);
      dn[i] = dot3(fr, df); dt1[i] = dot3(fr + 3, df); dt2[i] = dot3(fr + 6, df);
    }
    /* contact parameters: friction = elementwise max, solref / solimp mixed with equal weights */
    real mu = (real)(m->geom_friction[g1][0] > m->geom_friction[g2][0] ? m->geom_friction[g1][0] : m->geom_friction[g2][0]);
    double solref[2], solimp[5];
    for (int k = 0; k < 2; k++) solref[k] = 0.5 * (m->geom_solref[g1][k] + m->geom_solref[g2][k]);
    for (int k = 0; k < 5; k++) solimp[k] = 0.5 * (m->geom_solimp[g1][k] + m->geom_solimp[g2][k]);
    real w = (real)m->body_invweight0[b1] + (real)m->body_invweight0[b2];
    w = w + mu * mu * w;
    w = w * 2 * mu * mu / (real)m->impratio;
    for (int r = 0; r < 4; r++) {
      const real *tt = r < 2 ? dt1 : dt2;
      real sg = (r & 1) ? -mu : mu;
      for (int i = 0; i < m->nv; i++) J[i] = dn[i] + sg * tt[i];
      add_row(m, d, J, d->con_dist[c], w, solref, solimp);
    }
  }
}

/* ------------------------------------------------------------------ B.6 / BD.9 / BD.10 Newton solver */
typedef struct { real alpha, cost, d0, d1; } lspoint;
typedef struct {
  int nv, nefc;
  real qacc[MAXV], Ma[MAXV], Jaref[MAXEFC], grad[MAXV], Mgrad[MAXV], search[MAXV];
  real gauss, cost;
  int active[MAXEFC];
} sctx;

static void update_constraint(const omodel *m, odata *d, sctx *c) {
  real cost = 0;
  for (int r = 0; r < d->nefc; r++) {
    c->active[r] = c->Jaref[r] < 0;
    if (c->active[r]) cost += (real)0.5 * d->efc_D[r] * c->Jaref[r] * c->Jaref[r];
  }
  real g = 0;
  for (int i = 0; i < m->nv; i++) g += (c->Ma[i] - d->qfrc_smooth[i]) * (c->qacc[i] - d->qacc_smooth[i]);
  c->gauss = (real)0.5 * g;
  c->cost = cost + c->gauss;
}
static void ctx_create(const omodel *m, odata *d, sctx *c, const real *qacc) {
  for (int i = 0; i < m->nv; i++) c->qacc[i] = qacc[i];
  for (int r = 0; r < d->nefc; r++) {
    real s = 0;
    for (int i = 0; i < m->nv; i++) s += d->efc_J[r][i] * qacc[i];
    c->Jaref[r] = s - d->efc_aref[r];
  }
  mul_M(m->nv, d->M, c->Ma, qacc);
  update_constraint(m, d, c);
}
static void update_gradient(const omodel *m, odata *d, sctx *c) {
  int nv = m->nv;
  real H[MAXV][MAXV], HL[MAXV][MAXV];
  for (int i = 0; i < nv; i++) {
    real fc = 0;
    for (int r = 0; r < d->nefc; r++) if (c->active[r]) fc += d->efc_J[r][i] * (-d->efc_D[r] * c->Jaref[r]);
    c->grad[i] = c->Ma[i] - d->qfrc_smooth[i] - fc;
#ifdef ORACLE_COUNT
    for (int j = 0; j <= i; j++) {               /* H is symmetric: the op count books one triangle */
#else
    for (int j = 0; j < nv; j++) {
#endif
      real s = d->M[i][j];
      for (int r = 0; r < d->nefc; r++) if (c->active[r]) s += d->efc_J[r][i] * d->efc_D[r] * d->efc_J[r][j];
      H[i][j] = s;
#ifdef ORACLE_COUNT
      H[j][i] = s;
#endif
    }
  }
  chol(nv, H, HL);
  chol_solve(nv, HL, c->Mgrad, c->grad);
}
static lspoint ls_point(const odata *d, const sctx *c, real alpha, const real *jv, real quad[][3], const real *qg) {
  real q0 = qg[0], q1 = qg[1], q2 = qg[2];
  for (int r = 0; r < d->nefc; r++) if (c->Jaref[r] + alpha * jv[r] < 0) { q0 += quad[r][0]; q1 += quad[r][1]; q2 += quad[r][2]; }
  lspoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.d0 = 2 * alpha * q2 + q1;
  p.d1 = 2 * q2 + (q2 == 0 ? (real)MJ_MINVAL : (real)0);
  return p;
}
static int in_bracket(lspoint x, lspoint y) {
  return ((x.d0 < y.d0) && (y.d0 < 0)) || ((x.d0 > y.d0) && (y.d0 > 0));
}
static void linesearch(const omodel *m, odata *d, sctx *c) {
  int nv = m->nv;
  real sn = 0;
  for (int i = 0; i < nv; i++) sn += c->search[i] * c->search[i];
  real smag = RSQRT(sn) * (real)m->meaninertia * (nv > 1 ? nv : 1);
  real gtol = (real)m->tolerance * (real)m->ls_tolerance * smag;
  real mv[MAXV], jv[MAXEFC];
  real (*quad)[3] = malloc(sizeof(real) * 3 * (d->nefc + 1));
  mul_M(nv, d->M, mv, c->search);
  real qg[3] = {c->gauss, 0, 0};
  for (int i = 0; i < nv; i++) { qg[1] += c->search[i] * c->Ma[i] - c->search[i] * d->qfrc_smooth[i]; qg[2] += (real)0.5 * c->search[i] * mv[i]; }
  for (int r = 0; r < d->nefc; r++) {
    real s = 0;
    for (int i = 0; i < nv; i++) s += d->efc_J[r][i] * c->search[i];
    jv[r] = s;
    quad[r][0] = (real)0.5 * c->Jaref[r] * c->Jaref[r] * d->efc_D[r];
    quad[r][1] = jv[r] * c->Jaref[r] * d->efc_D[r];
    quad[r][2] = (real)0.5 * jv[r] * jv[r] * d->efc_D[r];
  }
  lspoint p0 = ls_point(d, c, 0, jv, quad, qg);
  lspoint lo = ls_point(d, c, p0.alpha - p0.d0 / p0.d1, jv, quad, qg), hi;
  if (lo.d0 < p0.d0) { hi = p0; } else { hi = lo; lo = p0; }
  int swap = 1;
  for (int it = 0; it < m->ls_iterations; it++) {
    int done = !swap;
    done |= (lo.d0 < 0) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0) && (hi.d0 < gtol);
    if (done) break;
    lspoint lo_next = ls_point(d, c, lo.alpha - lo.d0 / lo.d1, jv, quad, qg);
    lspoint hi_next = ls_point(d, c, hi.alpha - hi.d0 / hi.d1, jv, quad, qg);
    lspoint mid = ls_point(d, c, (real)0.5 * (lo.alpha + hi.alpha), jv, quad, qg);
    int s1 = in_bracket(lo, lo_next); if (s1) lo = lo_next;
    int s2 = in_bracket(lo, mid);     if (s2) lo = mid;
    int s3 = in_bracket(lo, hi_next); if (s3) lo = hi_next;
    int s4 = in_bracket(hi, hi_next); if (s4) hi = hi_next;
    int s5 = in_bracket(hi, mid);     if (s5) hi = mid;
    int s6 = in_bracket(hi, lo_next); if (s6) hi = lo_next;
    swap = s1 | s2 | s3 | s4 | s5 | s6;
#ifdef ORACLE_DEBUG
    printf("[ora]  it %d cand %.7g %.7g %.7g -> lo a %.7g d0 %.5g  hi a %.7g d0 %.5g swaps %d%d%d%d%d%d\n", it, (double)lo_next.alpha, (double)hi_next.alpha, (double)mid.alpha, (double)lo.alpha, (double)lo.d0, (double)hi.alpha, (double)hi.d0, s1, s2, s3, s4, s5, s6);
#endif
  }
  int improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
#ifdef ORACLE_DEBUG
  printf("[ora] p0 cost %.9g d0 %.6g d1 %.6g | lo a %.7g cost %.9g d0 %.6g | hi a %.7g cost %.9g d0 %.6g | gtol %.3g\n", (double)p0.cost, (double)p0.d0, (double)p0.d1, (double)lo.alpha, (double)lo.cost, (double)lo.d0, (double)hi.alpha, (double)hi.cost, (double)hi.d0, (double)gtol);
#endif
  real alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (improved) {
    for (int i = 0; i < nv; i++) { c->qacc[i] += c->search[i] * alpha; c->Ma[i] += mv[i] * alpha; }
    for (int r = 0; r < d->nefc; r++) c->Jaref[r] += jv[r] * alpha;
  }
  free(quad);
}
static void solve(const omodel *m, odata *d) {
  int nv = m->nv;
  if (d->nefc == 0) { for (int i = 0; i < nv; i++) d->qacc[i] = d->qacc_smooth[i]; memcpy(d->qacc_warmstart, d->qacc, sizeof(real) * nv); return; }
  sctx *c = malloc(sizeof(sctx)), *c2 = malloc(sizeof(sctx));
  ctx_create(m, d, c, d->qacc_warmstart);
  ctx_create(m, d, c2, d->qacc_smooth);
  const real *start = c->cost < c2->cost ? d->qacc_warmstart : d->qacc_smooth;
#ifdef ORACLE_DEBUG
  printf("[ora] nefc %d cost_w %.9g cost_s %.9g use_warm %d\n", d->nefc, (double)c->cost, (double)c2->cost, (int)(c->cost < c2->cost));
#endif
  real q0[MAXV];
  memcpy(q0, start, sizeof(real) * nv);
  ctx_create(m, d, c, q0);
  update_gradient(m, d, c);
  for (int i = 0; i < nv; i++) c->search[i] = -c->Mgrad[i];
  for (int it = 0; it < m->iterations; it++) {
    real prev = c->cost;
    linesearch(m, d, c);
    update_constraint(m, d, c);
    update_gradient(m, d, c);
    for (int i = 0; i < nv; i++) c->search[i] = -c->Mgrad[i];
    if (m->iterations > 1) {
      /* MJX while_loop termination (not reached with iterations = 1) */
      real scale = 1 / ((real)m->meaninertia * (nv > 1 ? nv : 1));
      real gn = 0; for (int i = 0; i < nv; i++) gn += c->grad[i] * c->grad[i];
      if ((prev - c->cost) * scale < (real)m->tolerance || RSQRT(gn) * scale < (real)m->tolerance) break;
    }
  }
  for (int i = 0; i < nv; i++) { d->qacc[i] = c->qacc[i]; d->qacc_warmstart[i] = c->qacc[i]; }
  free(c); free(c2);
}

/* ------------------------------------------------------------------ B.1 forward, B.8 euler */
static void forward(const omodel *m, odata *d) {
  OSTAGE(0);                   /* stages of the op counter (count_real.h); no-ops in the normal builds */
  kinematics(m, d);
  OSTAGE(1);
  com_pos(m, d);
  crb(m, d);
  chol(m->nv, d->M, d->L);
  OSTAGE(3);
  collision(m, d);
  OSTAGE(2);
  com_vel(m, d);
  passive(m, d);
  rne(m, d);
  for (int i = 0; i < m->nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i];
  chol_solve(m->nv, d->L, d->qacc_smooth, d->qfrc_smooth);
  OSTAGE(4);
  make_constraint(m, d);       /* row values do not depend on the velocity stage order */
  OSTAGE(5);
  solve(m, d);
  OSTAGE(7);
}
static void euler(const omodel *m, odata *d) {
  real dt = (real)m->timestep;
  for (int i = 0; i < m->nv; i++) d->qvel[i] += dt * d->qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == J_HINGE) { d->qpos[qa] += dt * d->qvel[da]; continue; }
    for (int k = 0; k < 3; k++) d->qpos[qa+k] += dt * d->qvel[da+k];
    real v[3] = {d->qvel[da+3], d->qvel[da+4], d->qvel[da+5]};
    real nrm = normalize3(v), ang = dt * nrm, s = RSIN(ang * (real)0.5);
    real qr[4] = {RCOS(ang * (real)0.5), s * v[0], s * v[1], s * v[2]}, qn[4];
    quat_mul(qn, d->qpos + qa + 3, qr);
    real n = RSQRT(qn[0]*qn[0]+qn[1]*qn[1]+qn[2]*qn[2]+qn[3]*qn[3]);
    for (int k = 0; k < 4; k++) d->qpos[qa+3+k] = qn[k] / n;
  }
}

/* ================================================================== exported entry points */
#ifdef __cplusplus
extern "C" {
#endif
#ifdef ORACLE_COUNT
static long long g_total[ORACLE_NSTAGE];
static void count_flush(void) {
  #pragma omp critical
  for (int k = 0; k < ORACLE_NSTAGE; k++) { g_total[k] += g_cnt[k]; g_cnt[k] = 0; }
}
/* stage totals since the last call: 0 kinematics, 1 com_pos + CRBA + factor, 2 velocity / passive / RNE / qacc_smooth,
 * 3 narrow phase, 4 constraint rows, 5 Newton + line search, 6 Euler, 7 not counted (I/O conversion, dense unobservable work) */
int oracle_count_read(long long *out) {
  for (int k = 0; k < ORACLE_NSTAGE; k++) { out[k] = g_total[k]; g_total[k] = 0; }
  return ORACLE_NSTAGE;
}
#else
static void count_flush(void) {}
#endif

/* One `mjx.forward` at (qpos, qvel) with the given warm start; returns qacc and a few intermediates
 * for unit tests (any output pointer may be NULL). */
int oracle_forward(const omodel *m, const double *qpos, const double *qvel, const double *warm,
                   double *qacc, double *M_out, double *qfrc_bias, double *qfrc_passive, double *con_dist,
                   double *xpos, double *xquat, double *site_tcp, double *con_pos, double *con_frame, int *nefc) {
  odata *d = calloc(1, sizeof(odata));
  for (int i = 0; i < m->nq; i++) d->qpos[i] = (real)qpos[i];
  for (int i = 0; i < m->nv; i++) { d->qvel[i] = (real)qvel[i]; d->qacc_warmstart[i] = warm ? (real)warm[i] : (real)0; }
  forward(m, d);
  if (qacc) for (int i = 0; i < m->nv; i++) qacc[i] = d->qacc[i];
  if (M_out) for (int i = 0; i < m->nv; i++) for (int j = 0; j < m->nv; j++) M_out[i * m->nv + j] = d->M[i][j];
  if (qfrc_bias) for (int i = 0; i < m->nv; i++) qfrc_bias[i] = d->qfrc_bias[i];
  if (qfrc_passive) for (int i = 0; i < m->nv; i++) qfrc_passive[i] = d->qfrc_passive[i];
  if (con_dist) for (int c = 0; c < m->ncon; c++) con_dist[c] = d->con_dist[c];
  if (con_pos) for (int c = 0; c < m->ncon; c++) for (int k = 0; k < 3; k++) con_pos[3*c+k] = d->con_pos[c][k];
  if (con_frame) for (int c = 0; c < m->ncon; c++) for (int k = 0; k < 9; k++) con_frame[9*c+k] = d->con_frame[c][k];
  if (xpos) for (int b = 0; b < m->nbody; b++) for (int k = 0; k < 3; k++) xpos[3*b+k] = d->xpos[b][k];
  if (xquat) for (int b = 0; b < m->nbody; b++) for (int k = 0; k < 4; k++) xquat[4*b+k] = d->xquat[b][k];
  if (site_tcp) for (int k = 0; k < 3; k++) site_tcp[k] = d->site_tcp[k];
  if (nefc) *nefc = d->nefc;
  free(d);
  count_flush();
  return 0;
}

/* compute_rollout_batch (mjx_planner.py:266-274, vmapped at :123):
 *   thetadot [B, ndof*T] (index dof*T + t), q0/v0 [ndof]; every sample starts from the snapshot
 *   (qpos_init, qvel_init, warm_init) with qpos[:ndof]=q0, qvel[:ndof]=v0.
 *   Outputs: theta [B, ndof*T] (post-step qpos, dof-major), eef_pos [B,T,3], eef_rot [B,T,4]
 *   (pre-step), collision [B,T,nrobot] (pre-step dist of the robot-involving slots, slot order).
 *   Optional full-state outputs for the tests: qpos_out [B,T,nq] post-step, qacc_out [B,T,nv]. */
int oracle_rollout(const omodel *m, int B, int T, int ndof, const double *thetadot, const double *q0, const double *v0,
                   const double *qpos_init, const double *qvel_init, const double *warm_init,
                   double *theta, double *eef_pos, double *eef_rot, double *collision,
                   double *qpos_out, double *qacc_out, int nthreads) {
  int nrobot = 0;
  for (int c = 0; c < m->ncon; c++) nrobot += m->slot_robot[c] != 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  #pragma omp parallel
  {
    odata *d = malloc(sizeof(odata));
    #pragma omp for schedule(dynamic, 4)
    for (int s = 0; s < B; s++) {
      memset(d, 0, sizeof(odata));
      for (int i = 0; i < m->nq; i++) d->qpos[i] = (real)qpos_init[i];
      for (int i = 0; i < m->nv; i++) { d->qvel[i] = (real)qvel_init[i]; d->qacc_warmstart[i] = (real)warm_init[i]; }
      for (int i = 0; i < ndof; i++) { d->qpos[i] = (real)q0[i]; d->qvel[i] = (real)v0[i]; }
      for (int t = 0; t < T; t++) {
        for (int i = 0; i < ndof; i++) d->qvel[i] = (real)thetadot[(size_t)s * ndof * T + i * T + t];   /* mjx_planner.py:254 */
        forward(m, d);
        for (int k = 0; k < 3; k++) eef_pos[((size_t)s * T + t) * 3 + k] = d->site_tcp[k];
        for (int k = 0; k < 4; k++) eef_rot[((size_t)s * T + t) * 4 + k] = d->xquat[m->hande_body][k];
        if (collision) { int o = 0; for (int c = 0; c < m->ncon; c++) if (m->slot_robot[c]) collision[((size_t)s * T + t) * nrobot + o++] = d->con_dist[c]; }
        if (qacc_out) for (int i = 0; i < m->nv; i++) qacc_out[((size_t)s * T + t) * m->nv + i] = d->qacc[i];
        OSTAGE(6);
        euler(m, d);
        OSTAGE(7);
        for (int i = 0; i < ndof; i++) theta[(size_t)s * ndof * T + i * T + t] = d->qpos[i];
        if (qpos_out) for (int i = 0; i < m->nq; i++) qpos_out[((size_t)s * T + t) * m->nq + i] = d->qpos[i];
      }
    }
    free(d);
    count_flush();
  }
  return 0;
}

int oracle_sizeof_model(void) { return (int)sizeof(omodel); }
int oracle_real_bytes(void) { return (int)sizeof(real); }

/* ------------------------------------------------------------------ isolated colliders for unit tests
 * type: 0 plane-capsule, 1 capsule-capsule, 2 capsule-box, 3 plane-box, 4 box-box.
 * pos/mat: world pose (mat row-major), size: mujoco size triple.  Outputs up to 4 slots. */
int oracle_collide(int type, const double *p1, const double *m1, const double *s1, const double *p2, const double *m2,
                   const double *s2, double *dist, double *pos, double *frame) {
  real P1[3], M1[9], S1[3], P2[3], M2[9], S2[3], d[4] = {1, 1, 1, 1}, ps[4][3], fr[4][9];
  memset(ps, 0, sizeof ps); memset(fr, 0, sizeof fr);
  for (int k = 0; k < 3; k++) { P1[k] = (real)p1[k]; P2[k] = (real)p2[k]; S1[k] = (real)s1[k]; S2[k] = (real)s2[k]; }
  for (int k = 0; k < 9; k++) { M1[k] = (real)m1[k]; M2[k] = (real)m2[k]; }
  int n = 0;
  if (type == 0) { plane_capsule(P1, M1, P2, M2, S2, d, ps, fr); n = 2; }
  else if (type == 1) { capsule_capsule(P1, M1, S1, P2, M2, S2, d, ps, fr); n = 1; }
  else if (type == 2) { capsule_box(P1, M1, S1, P2, M2, S2, d, ps, fr, 1); n = 2; }
  else if (type == 5) { capsule_box(P1, M1, S1, P2, M2, S2, d, ps, fr, 0); n = 2; }   /* round-1 restatement */
  else if (type == 3) { plane_box(P1, M1, P2, M2, S2, d, ps, fr); n = 4; }
  else if (type == 4) { box_box(P1, M1, S1, P2, M2, S2, d, ps, fr); n = 4; }
  for (int i = 0; i < n; i++) {
    dist[i] = d[i];
    for (int k = 0; k < 3; k++) pos[3*i+k] = ps[i][k];
    for (int k = 0; k < 9; k++) frame[9*i+k] = fr[i][k];
  }
  return n;
}
#ifdef __cplusplus
}
#endif
