// upright_b200 — the batched MPC solve kernel: one warp per MPC instance.
//
// Replaces, for B independent instances, the body of
// `ControllerInterface::advanceMpc()` (upright_control/src/pybindings.cpp:376,
// solver constructed at upright_control/src/controller_interface.cpp:395-398):
//   per SQP iteration:  linearise all knots (lane = d/dx_j tangent direction)
//                       -> OCP-QP by a primal-dual interior point method whose
//                          Newton systems are Riccati recursions held in shared
//                          memory (structured A = A3 (x) I, B = B3 (x) I)
//                       -> filter line search on the true nonlinear functions
//                          (lane = knot).
// The algorithm (row families, soft/hard handling, constants) is specified in
// DESIGN.md §4 and is the same one oracle/oracle.cpp implements in fp64 with
// dense linear algebra.
//
// Round 2: REDUCED STAGE + MIXED PRECISION (oracle/reduced_lab.h is the CPU study of the same
// arithmetic).  The contact forces are eliminated from every stage in range-space form before the Riccati
// recursion, which then runs on [jerk; state] only (nq + 3 nq variables: 36 for every Thing configuration):
//     D      = force block of the stage matrix (diagonal, or 3 x 3 per contact with friction pyramids)
//     S      = R^-1 + Df D^-1 Df'                  (6 nb x 6 nb; per body unless bodies share contacts)
//     M_xx  += C' S^-1 C = G'G,  G = L^-1 C        (what the penalty term leaves once the forces are gone)
//     m_x   += G' L^-1 (e + R^-1 y - Df D^-1 m_f)
//     lambda = S^-1 (e + R^-1 y - Df D^-1 m_f + C dx),   df = -D^-1 (m_f + Df' lambda)
// which is the exact block elimination of df (Woodbury on M_ff = D + Df' R Df): no penalty weight enters a
// matrix that is factorised, only its reciprocal.  Two arithmetic types: F (float in the product kernels) for the
// Riccati matrices, factors and directions; R = double for the iterate, the slack / multiplier records, every
// residual and gradient, the linearisation and the small force block.  The fp64 validation kernels use F = double.
#pragma once
#include <math_constants.h>

#include <type_traits>

#include "ub_device.cuh"

namespace ub {

// Per-problem workspace layout in units of F.  One constexpr function serves the host (make_layout in
// ub_api.cu) and the kernels specialised on compile-time dimensions, where every offset becomes an
// immediate of the load/store instruction.  Blocks marked (R) hold doubles: rw = sizeof(double) / sizeof(F)
// units per element, 16-byte aligned like every block.
struct Layout {
    int Z, DZ, GAP, LG, LC, LR, LJP, LHO, LJO, DF, RHOE, YE, RHOT, YT, TL, DD, GP, VE, FAC, WF, FBB;
    int XN, UN, XW, UW, TG, BD, LIA, LJA, XO, DXO;
    int TGQ, LRO, LJQ;   // end-effector orientation cost: target quaternions [N+1, 4], error (R) [N+1, 3], Jacobian [N+1, 3, nq]
    int total;
    // force bundle of one stage (workspace FBB + k * bsize, and the same layout in shared memory at sFB):
    // G = L^-1 C [neq][nx] | L^-1 per group (double) | D^-1 per contact (double) | g_lambda (double) | q (double)
    int bG, bL, bD, bGl, bQ, bsize;
    // shared memory (units of F, per warp)
    int sM, sP, sPv, sFB, sGf, sFv, sFl, sVec, sDst, sDxn, sRv, sScr, sDFC, sUS, sCst, sTL, sDD, sSmZ, sSmX, sSmU, sSmJ, sSmW;
    int sBar;   // two mbarriers (UB_TMA_STAGE variant)
    int s_total;
    int rw;   // units per double
};
// side records of a stage are staged in shared memory (cp.async, one stage ahead) up to this many rows (72 covers cfg5's
// 68-row stage: +18 % at 1024 instances; cfg3's 164-row stage loses more to the resident warps the buffers cost than it
// gains, profiles/r2_v6_staging_ab.txt)
// cycle counters of the phases (tools/gpu_phase_profile.py): compiled in only for the profiling variant
// (python tools/build_variant.py profile -DUB_PROFILE=1) — in the product build they cost registers and, once spilled, a
// local-memory read-modify-write per stage
#ifndef UB_PROFILE
#define UB_PROFILE 0
#endif
#if UB_PROFILE
#define UB_CLK() clock64()
#define UB_ACC(var, expr) (var) += (expr)
#else
#define UB_CLK() 0LL
#define UB_ACC(var, expr) ((void)0)
#endif
#ifndef UB_TMA_STAGE
#define UB_TMA_STAGE 0
#endif
#ifndef UB_STAGE_ROWS_MAX
#define UB_STAGE_ROWS_MAX 72
#endif
struct LayoutDims {
    int N, nq, nx, nu, neq, nfc, nterm, nrow, nobs, nb, nc, nf;
    int rw;          // sizeof(double) / sizeof(F)
    int ngrp, ng;    // force block: ngrp groups of ng equality rows (per body: nb x 6; bodies sharing contacts: 1 x 6 nb)
    int ori = 0;     // end-effector orientation cost enabled
    int nia = 0;     // rows of the inertial-alignment cost (0 or 2)
    int obsw = 0;    // width of the obstacle-family rows (0 -> nq)
    int nxo = 0;     // dynamic-obstacle states (9 per obstacle)
};
__host__ __device__ constexpr int ub_round4(int n) { return (n + 3) / 4 * 4; }
__host__ __device__ constexpr int ub_max(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr Layout compute_layout(const LayoutDims d) {
    Layout L{};
    const int N = d.N, nx = d.nx, nu = d.nu, nz = d.nu + d.nx, nq = d.nq, nr = d.nq + d.nx, rw = d.rw;
    const int fbd = d.nc * (d.nf == 3 ? 9 : 1), fbl = d.ngrp * d.ng * d.ng;
    const int fstride = ub_round4(nq * (nr + 1));
    L.rw = rw;
    int o = 0;
    L.Z = o;    o += ub_round4(rw * (N + 1) * nz);          // (R) QP iterate
    L.DZ = o;   o += ub_round4((N + 1) * nz);               // Newton direction [dj; df; dx]
    L.GAP = o;  o += ub_round4(rw * N * nx);                // (R) dynamics defect of the iterate
    L.LG = o;   o += ub_round4(rw * N * d.neq);             // (R) object-dynamics rows: value at the iterate
    L.LC = o;   o += ub_round4(N * d.neq * nx);             // their state Jacobians C [k][neq][nx]
    L.LR = o;   o += ub_round4(rw * (N + 1) * 3);           // (R) tool position
    L.LJP = o;  o += ub_round4((N + 1) * 3 * nq);
    L.LHO = o;  o += ub_round4((N + 1) * d.nobs);
    L.LJO = o;  o += ub_round4((N + 1) * d.nobs * (d.obsw > 0 ? d.obsw : nq));
    L.DF = o;   o += ub_round4(d.neq * d.nfc);              // d g / d f dense [neq][nfc] (constant over the solve)
    L.RHOE = o; o += ub_round4(rw * N * d.neq);             // (R) weights of the equality rows
    L.YE = o;   o += ub_round4(rw * N * d.neq);             // (R) their multipliers (hard rows)
    L.RHOT = o; o += ub_round4(rw * d.nterm);
    L.YT = o;   o += ub_round4(rw * d.nterm);
    L.TL = o;   o += ub_round4(rw * (N + 1) * d.nrow * 4);  // (R) side records {t_lo, t_hi, lam_lo, lam_hi}
    L.DD = o;   o += ub_round4((N + 1) * d.nrow * 4);       // their steps {dt_lo, dt_hi, dlam_lo, dlam_hi}
    L.GP = o;   o += ub_round4(rw * (N + 1) * nz);          // (R) predictor stage gradients kept for the corrector
    L.VE = o;   o += ub_round4(rw * N * d.neq);             // (R) e + y / rho of the equality rows
    L.FAC = o;  o += ub_round4(N * fstride);                // Riccati factor blocks [L; Y], column-major, column length nr + 1
    L.WF = o;   o += ub_round4(N * nq);
    {   // force bundle: one contiguous block per stage, fetched by one copy in the forward / corrector passes
        int b = 0;
        L.bG = b;  b += ub_round4(d.neq * nx);              // G = L^-1 C
        L.bL = b;  b += ub_round4(rw * fbl);                // (R) L^-1 of S per group
        L.bD = b;  b += ub_round4(rw * fbd);                // (R) D^-1 per contact
        L.bGl = b; b += ub_round4(rw * d.neq);              // (R) g_lambda = L^-1 (e + y / rho - Df D^-1 m_f)
        L.bQ = b;  b += ub_round4(rw * d.nfc);              // (R) q = D^-1 m_f
        L.bsize = b;
    }
    L.FBB = o;  o += N * L.bsize;
    L.XN = o;   o += ub_round4((N + 1) * nx);
    L.UN = o;   o += ub_round4(N * nu);
    // the iterate, the desired positions and the body parameters live in the instance workspace too, so that
    // the solve needs ONE per-instance base address (the batch arrays are touched at entry and exit only)
    L.XW = o;   o += ub_round4((N + 1) * nx);
    L.UW = o;   o += ub_round4(N * nu);
    L.TG = o;   o += ub_round4((N + 1) * 3);
    L.BD = o;   o += ub_round4((d.nb > 0 ? d.nb : 1) * 10);
    L.LIA = o;  o += ub_round4(N * d.nia);
    L.LJA = o;  o += ub_round4(N * d.nia * nx);
    L.XO = o;   o += ub_round4((N + 1) * d.nxo);
    L.DXO = o;  o += ub_round4((N + 1) * d.nxo);
    L.TGQ = o;  o += d.ori ? ub_round4((N + 1) * 4) : 0;
    L.LRO = o;  o += d.ori ? ub_round4(rw * (N + 1) * 3) : 0;
    L.LJQ = o;  o += d.ori ? ub_round4((N + 1) * 3 * nq) : 0;
    L.total = o;
    int s = 0;
    // (the stage-matrix buffer doubles as the scratch of the S build when bodies share contacts: D^-1 N' per contact
    // and side, [c][side][6][nf] doubles — the force block is eliminated before the matrix is built)
    const int us = (d.ngrp == 1 && d.nb > 1) ? rw * d.nc * 2 * 6 * d.nf : 0;
    L.sM = s;   s += ub_round4(ub_max(ub_max(nr * (nr | 1), 2 * fstride), us));
    L.sP = s;   s += ub_round4(nx * nx);
    L.sFB = s;  s += L.bsize;                               // force bundle of the stage (C rows first, G in place)
    L.sGf = s;  s += ub_round4(d.neq);                      // g_lambda in F for the Schur update
    L.sFv = s;  s += ub_round4(rw * d.neq);                 // (R) v = e + y / rho, then rhs
    L.sFl = s;  s += ub_round4(rw * d.neq);                 // (R) lambda / scratch
    L.sVec = s; s += ub_round4(rw * nz);                    // (R) stage gradient [j; f; x]
    // the forward sweeps (B, D) and the backward sweeps (A, C) never run at the same time: their vectors share storage
    L.sDst = s; L.sRv = s; s += ub_round4(ub_max(nz, nr + 1));   // stage direction [dj; df; dx] | Riccati right-hand side [m_j; m_x]
    L.sDxn = s; L.sPv = s; s += ub_round4(nx);                   // next state direction | cost-to-go gradient
    L.sScr = s; s += ub_round4(rw * ub_max(ub_max(d.neq, d.nobs), ub_max(d.nterm, 16)));   // (R) per-row scratch
    L.sDFC = s; s += ub_round4(d.nc * 2 * 6 * d.nf);        // d g / d f per contact and side: [c][side][6][nf] (constant over the solve)
    L.sUS = L.sM;
    const bool staged = d.nrow <= UB_STAGE_ROWS_MAX;
    L.sCst = s; s += staged ? ub_round4(d.neq * nx) : 0;    // C rows of the stage a factor pass visits next
    L.sTL = s;  s += staged ? ub_round4(rw * d.nrow * 4) : 0;
    L.sDD = s;  s += staged ? ub_round4(d.nrow * 4) : 0;
    L.sSmZ = s; s += staged ? ub_round4(rw * nz) : 0;       // (R) staged z_k (or predictor gradient)
    L.sSmX = s; s += staged ? ub_round4(nx) : 0;
    L.sSmU = s; s += staged ? ub_round4(nu) : 0;
    L.sSmJ = s; s += staged ? ub_round4(3 * nq) : 0;
    L.sSmW = s; s += ub_round4(nq);
    L.sBar = s; s += UB_TMA_STAGE ? 4 : 0;
    L.s_total = s;
    return L;
}

template <typename F>
struct BatchArgs {
    const F* x0;      // [B, nx]
    const F* target;  // [B, N+1, 3]
    const F* body;    // [B, nb, 10] or null
    F* X;             // [B, N+1, nx]  solution out (device memory, or mapped pinned host memory: the host path lets
    F* U;             // [B, N, nu]    the kernel write results home while other instances are still being solved)
    const F* Xin;     // warm start in (UB_WARM_START); may alias X
    const F* Uin;
    F* K;             // [B, N, nu, nx] or null
    int32_t* status;  // [B]
    F* stats;         // [B, UB_STATS] or null
    F* ws;            // [slots, layout.total]
    int B;
    int warm;
    int stop_after;   // debug: 0 = full solve, 1 = stop after first linearisation, 2 = after first QP
    int gain_stages;  // K holds the gains of stages [0, gain_stages) only: [B, gain_stages, nu, nx]
    // Work queue: with `queue` set the grid is persistent (one workspace slot per resident warp) and every warp
    // keeps taking the next unsolved instance from the counter, so a warp whose instance converged early does
    // not idle until the slowest instance of its CTA is done.  Without it warp w of CTA c solves instance
    // c * warps_per_cta + w in workspace slot of the same index (test aid: intermediate blocks are inspected).
    int* queue;
    int n_slots;      // workspace slots (= warps that may work); B in the static mode
    int nxt;          // columns of x0 / X / Xin: robot state + dynamic-obstacle states
    int tstride;      // columns of a target row: 3, or 7 with the desired quaternion [x y z w]
    int align;        // phase-align the warps of a CTA at every interior-point iteration: warps per group, 0 = off (CtaAlign)
    // Multi-GPU gather fused into the solve: besides X / U of this rank, every solved instance is
    // stored straight into the gathered buffers of the peer GPUs (peer-mapped pointers, NVLink P2P stores from the
    // epilogue) at row gather_row + b — the all-gather happens instance by instance while the rest of the batch is
    // still being solved, with no collective kernel and no copy afterwards.
    F* Xg[UB_MAX_GATHER];
    F* Ug[UB_MAX_GATHER];
    int ngather;
    long long gather_row;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <typename T>
__device__ __forceinline__ T tinf() { return T(1e30); }

template <typename T>
struct Perf {
    T cost, dyn, eq, ineq, max_eq, min_margin;
    __device__ __forceinline__ T violation() const { return sqrt(dyn + eq + ineq); }
};

// Compile-time problem dimensions (the BASELINE configurations) ...
template <int NQ_, int NF_, int NC_, int NB_, int NOBS_ = 0, int N_ = 20, bool COUPLED_ = false>
struct StaticDims {
    static constexpr bool kStatic = true;
    static constexpr int nq = NQ_, nf = NF_, nc = NC_, nb = NB_, nobs = NOBS_, N = N_;
    static constexpr int nx = 3 * NQ_, nfc = NF_ * NC_, nu = NQ_ + NF_ * NC_, nz = nu + nx, nr = NQ_ + nx;
    static constexpr int neq = 6 * NB_, nfric = (NF_ == 3) ? 5 * NC_ : 0;
    static constexpr int nbox_u = nfc > 0 ? nu : nq;
    static constexpr int nterm = 3 + 2 * NQ_, nrow = nbox_u + nx + nfric + NOBS_;
    static constexpr int ngrp = COUPLED_ ? 1 : NB_, ng = COUPLED_ ? 6 * NB_ : 6;
    template <typename F>
    __host__ __device__ static constexpr Layout layout() {
        return compute_layout(LayoutDims{N, nq, nx, nu, neq, nfc, nterm, nrow, nobs, nb, nc, nf, int(sizeof(double) / sizeof(F)), ngrp, ng});
    }
};
// ... or only the robot compile-time (nq = 6: fixed-base UR10, 9: Thing) and everything else at run time.  The
// Riccati recursion depends on nq alone, so every configuration takes the same register-blocked factorisation.
template <int NQ_>
struct RuntimeDims {
    static constexpr bool kStatic = false;
    static constexpr int nq = NQ_, nx = 3 * NQ_, nr = NQ_ + nx;
    static constexpr int nf = 0, nc = 0, nb = 0, nfc = 0, nu = 0, nz = 0, neq = 0, nfric = 0, nbox_u = 0;
    static constexpr int nobs = 0, N = 0, nterm = 3 + 2 * NQ_, nrow = 0, ngrp = 0, ng = 0;
    template <typename F>
    __host__ __device__ static constexpr Layout layout() { return Layout{}; }
};

// pinf * ratio^(iterations left) > tol: the residual of the hard equality rows cannot reach the tolerance before the
// iteration cap at its present linear rate.  Out of line: two double logarithms that would otherwise sit (and take
// registers) in the interior-point loop of every kernel.
static __device__ __noinline__ bool hopeless_rate(double pinf, double pinf_prev, double tol, int it, int iter_max) {
    const double ratio = pinf / pinf_prev;
    return ratio >= 1.0 || double(it) + log(tol / pinf) / log(ratio) > double(iter_max);
}

// Phase alignment of the warps of a CTA: the interior-point loop is ~200 KB of SASS against a
// 32 KB instruction cache, and warps that walk it at the same time share what is fetched — measured 1.36x between a
// batch whose instances all take the same number of iterations and a mixed one (profiles/r2_v8_alignment_probe.txt).
// Every warp therefore meets the others of its CTA at the top of each interior-point iteration (a software barrier in
// shared memory: 16 warps, 7 meetings per solve).  Warps between two solves (line search, output, next linearisation)
// arrive when they reach their next first iteration; warps out of work keep attending until all are.  A meeting that
// does not complete within the spin limit switches the alignment off for the rest of the launch (never a hang).
struct CtaAlign {
    // one packed word per alignment group (up to four groups per CTA; group = warps with the same index modulo the
    // number of groups, i.e. warps of the same SM sub-partition):
    //   bits 0-5 arrivals of the current meeting | 6-11 warps of the group that still take instances | 12-17 warps of the
    //   group | bit 18 "all out of work" (published by the last arriver) | bit 19 alignment switched off | 20-31 meeting
    //   counter
    static constexpr unsigned kDone = 1u << 18, kOff = 1u << 19;
    unsigned word[4];
};

template <typename F, typename D>
struct Solver {
    using R = double;
    unsigned* al = nullptr;   // packed word of this warp's alignment group; null: no alignment
    // returns true once every warp of the group is out of work (or the alignment has been switched off)
    __device__ __forceinline__ bool align_wait() const {
        if (al == nullptr) return true;
        __syncwarp();
        unsigned w = 0;
        if (lane == 0) {
            volatile unsigned* vw = al;
            w = *vw;
            if ((w & CtaAlign::kOff) == 0) {
                const unsigned old = atomicAdd(al, 1u);
                const unsigned g = old >> 20;
                if ((old & 0x3fu) + 1u >= ((old >> 12) & 0x3fu)) {
                    // last arriver: nobody else touches the word now (every other warp of the group waits)
                    const unsigned working = (old >> 6) & 0x3fu;
                    w = (((g + 1u) & 0xfffu) << 20) | (old & (CtaAlign::kOff | (0x3fu << 12) | (0x3fu << 6))) |
                        (working == 0u ? CtaAlign::kDone : 0u);
                    __threadfence_block();
                    *vw = w;
                } else {
                    int spins = 0;
                    for (;;) {
                        w = *vw;
                        if ((w >> 20) != g || (w & CtaAlign::kOff)) break;
                        __nanosleep(40);
                        if (++spins > (1 << 20)) {   // ~0.1 s: give the alignment up, never hang
                            atomicOr(al, CtaAlign::kOff);
                            w = *vw;
                            break;
                        }
                    }
                }
                __threadfence_block();
            }
        }
        w = __shfl_sync(FULL, w, 0);
        return (w & (CtaAlign::kDone | CtaAlign::kOff)) != 0u;
    }
    static constexpr int kTS = WARP;
    static constexpr int RW = int(sizeof(R) / sizeof(F));
    const DevProblem<F>& P;    // shared-memory copy: arrays indexed per lane
    const DevProblem<F>& C;    // kernel-parameter copy (constant bank): scalars and uniformly indexed entries
    const DevProblem<R>& PR;   // the same constants in double (shared memory) for residuals and the linearisation
    const Layout& L;
    const int lane;
#define UB_DIM(FN, name) \
    __device__ __forceinline__ int FN() const { if constexpr (D::kStatic) return D::name; else return P.name; }
    UB_DIM(NU, nu) UB_DIM(NZ, nz) UB_DIM(NFC, nfc) UB_DIM(NEQ, neq) UB_DIM(NFRIC, nfric)
    UB_DIM(NBOXU, nbox_u) UB_DIM(NB, nb) UB_DIM(NC, nc) UB_DIM(NF, nf) UB_DIM(NN, N) UB_DIM(NOBS, nobs)
    UB_DIM(NGRP, ngrp) UB_DIM(NG, ng)
#undef UB_DIM
    static constexpr int kNQ = D::nq, kNX = D::nx, kNR = D::nr;
    __device__ __forceinline__ static constexpr int NQ() { return D::nq; }
    __device__ __forceinline__ static constexpr int NX() { return D::nx; }
    __device__ __forceinline__ static constexpr int NR() { return D::nr; }
    // end-effector box rows exist only in the run-time-dimension kernel (the specialised ones are not dispatched
    // for such problems)
    __device__ __forceinline__ bool EEBOX() const { if constexpr (D::kStatic) return false; else return P.eebox != 0; }
    // width of the dense rows of the obstacle family: the configuration (nq) for distances and the end-effector
    // box, the whole state when inertial-alignment constraint rows (which see v and a) ride along
    __device__ __forceinline__ int OBSW() const { if constexpr (D::kStatic) return D::nq; else return P.obsw; }
    __device__ __forceinline__ bool IACON() const { if constexpr (D::kStatic) return false; else return P.iacon != 0; }
    // projectile-path rows (run-time-dimension kernel only), last of the obstacle family
    __device__ __forceinline__ int NPROJ() const { if constexpr (D::kStatic) return 0; else return P.nproj; }
    __device__ __forceinline__ bool SPHERES() const { return NPAIRS() > 0 || NPROJ() > 0; }
    // dynamic obstacles (run-time-dimension kernel only): number of appended states
    __device__ __forceinline__ int NXO() const { if constexpr (D::kStatic) return 0; else return P.nxo; }
    // Sphere centres of the dynamic obstacles at knot k for the obstacle iterate XO + ao * DXO.  The obstacle states
    // are uncontrolled (system_dynamics.h:28-38): their Newton step DXO = exact rollout - iterate is known before
    // the QP, so they never enter it — their effect is the shift of the distance-row constants in linearize().
    template <typename T>
    __device__ __forceinline__ void place_dynamic_spheres(int k, T ao, T* sph) const {
        for (int s = 0; s < P.nsph; ++s)
            if (P.slink[s] <= -2) {
                const int o = (k * P.ndyn + (-2 - P.slink[s])) * 9;
                for (int c = 0; c < 3; ++c)
                    sph[3 * s + c] = T(ws[oXO() + o + c]) + (ao != T(0) ? ao * T(ws[oDXO() + o + c]) : T(0));
            }
    }
    // the inertial-alignment cost likewise (run-time-dimension kernel only)
    __device__ __forceinline__ bool IALIGN() const { if constexpr (D::kStatic) return false; else return P.iacost != 0; }
    // end-effector orientation cost (run-time-dimension kernel only; zero weight in every shipped configuration)
    __device__ __forceinline__ bool ORI() const { if constexpr (D::kStatic) return false; else return P.ori != 0; }
    __device__ __forceinline__ int NPAIRS() const { if constexpr (D::kStatic) return D::nobs; else return P.npairs; }
    // workspace / shared-memory offsets: immediates for the specialised kernels
#define UB_OFF(name) \
    __device__ __forceinline__ int o##name() const { if constexpr (D::kStatic) { constexpr Layout l = D::template layout<F>(); return l.name; } else return L.name; }
    UB_OFF(Z) UB_OFF(DZ) UB_OFF(GAP) UB_OFF(LG) UB_OFF(LC) UB_OFF(LR) UB_OFF(LJP) UB_OFF(LHO) UB_OFF(LJO) UB_OFF(DF)
    UB_OFF(RHOE) UB_OFF(YE) UB_OFF(RHOT) UB_OFF(YT) UB_OFF(TL) UB_OFF(DD) UB_OFF(GP) UB_OFF(VE) UB_OFF(FAC) UB_OFF(WF)
    UB_OFF(FBB) UB_OFF(bG) UB_OFF(bL) UB_OFF(bD) UB_OFF(bGl) UB_OFF(bQ) UB_OFF(bsize) UB_OFF(XN) UB_OFF(UN)
    UB_OFF(XW) UB_OFF(UW) UB_OFF(TG) UB_OFF(BD) UB_OFF(LIA) UB_OFF(LJA) UB_OFF(XO) UB_OFF(DXO) UB_OFF(TGQ) UB_OFF(LRO) UB_OFF(LJQ)
#undef UB_OFF
    __device__ __forceinline__ static constexpr int LDM() { return D::nr | 1; }
    __device__ __forceinline__ int NROW() const { return NBOXU() + NX() + NFRIC() + NOBS(); }
    __device__ __forceinline__ static constexpr int NTERM() { return 3 + 2 * D::nq; }
    __device__ __forceinline__ static constexpr int FSTRIDE() { return ub_round4(D::nq * (D::nr + 1)); }
    __device__ __forceinline__ int FBDN() const { return NC() * (NF() == 3 ? 9 : 1); }
    __device__ __forceinline__ int FBLN() const { return NGRP() * NG() * NG(); }
    // entry (i, j) of a stored Riccati factor block [L; Y]: column-major, column length nr + 1 — written straight
    // from the panel registers with coalesced stores, read conflict-free by every sweep
    __device__ __forceinline__ static constexpr int fidx(int i, int j) { return j * (D::nr + 1) + i; }
    // batch data of this instance
    F* ws;   // the one per-instance base address; X, U, target, body are instance-local blocks of it
    F* X;
    F* U;
    const F* target;
    const F* body;
    // shared memory of this warp
    F* sM;
    F* sP;
    F* sPv;
    F* sFB;   // force bundle of the current stage
    F* sC;    //   G (C rows on entry of the factor pass), dense [neq][nx]
    R* sS;    //   L^-1 per group
    R* sD;    //   D^-1 per contact
    R* sFg;   //   g_lambda
    R* sFq;   //   q
    F* sGf;
    R* sFv;
    R* sFl;
    F* sDFC;  // d g / d f per contact and side (constant over the solve)
    R* sUS;   // scratch of the S build when bodies share contacts
    F* sCst;  // C rows of the current stage in the factor pass of the staged kernels (cC points here or at sC)
    F* cC;
    R* sVec;
    F* sDst;
    F* sDxn;
    F* sRv;
    R* sScr;
    R* sTL;
    F* sDD;
    R* sSmZ;
    F* sSmX;
    F* sSmU;
    F* sSmJ;
    F* sSmW;
    template <typename T>
    __device__ __forceinline__ T* wsr(int off) const { return reinterpret_cast<T*>(ws + off); }
    __device__ __forceinline__ F* bundle(int k) const { return ws + oFBB() + k * obsize(); }
    // reciprocal in the residual arithmetic: exact in the fp64 validation kernels
    __device__ __forceinline__ static R rinv(R d) {
        if constexpr (std::is_same<F, R>::value) return R(1) / d;
        else return fast_rcp(d);
    }

    __device__ __forceinline__ void tsync() const { __syncwarp(); }
    template <typename T>
    __device__ __forceinline__ T tsum(T v) const { return warp_sum(v); }
    template <typename T>
    __device__ __forceinline__ T tmax(T v) const { return warp_max(v); }
    template <typename T>
    __device__ __forceinline__ T tmin(T v) const { return warp_min(v); }
    int nan_reason = 0;  // where a solve first went non-finite: 1 factorisation, 2 step, 4 complementarity, 5 step size, 3 iterate, 6 force block
    // phase cycle counters (profile mode, option stop_after = 9): linearise, factor pass, forward predictor, corrector
    // passes, line search; and inside the factor pass: gradient, matrix build, force block, factorisation
    long long t_lin = 0, t_fac = 0, t_swp = 0, t_side = 0, t_ls = 0;
    long long t_g = 0, t_f1 = 0, t_f2 = 0, t_f3 = 0;
    long long t_al = 0, t_qi = 0;   // waiting at the alignment meetings; set-up of the interior-point iteration

    __device__ Solver(const DevProblem<F>& P_, const DevProblem<F>& C_, const DevProblem<R>& PR_, const Layout& L_, int lane_)
        : P(P_), C(C_), PR(PR_), L(L_), lane(lane_) {}

    // ------------------------------------------------------------ row model
    // Inequality rows of a stage, in this order:
    //   [0, nbox_u)                    input box  (k < N)     controller_interface.cpp:165-169,330-356
    //   [nbox_u, +nx)                  state box  (k >= 1)    controller_interface.cpp:157-163
    //   [.., +nfric)                   friction   (k < N)     contact_constraints.h:49-77
    //   [.., +nobs)                    obstacles  (1<=k<N)    controller_interface.cpp:450-481
    __device__ __forceinline__ int row_family(int r) const {
        if (r < NBOXU()) return 0;
        if (r < NBOXU() + NX()) return 1;
        if (r < NBOXU() + NX() + NFRIC()) return 2;
        return 3;
    }
    __device__ __forceinline__ bool row_valid(int k, int fam) const {
        switch (fam) {
            case 0: return k < NN();
            case 1: return k >= 1;
            case 2: return k < NN();
            default: return k >= 1 && k < NN();
        }
    }
    __device__ __forceinline__ bool row_soft(int fam) const {
        return fam == 0 ? C.soft_u : (fam == 1 ? C.soft_x : C.soft_poly);
    }
    __device__ __forceinline__ R row_eps(int fam) const { return row_soft(fam) ? PR.invZ : PR.eps_hard; }
    // friction row coefficients on the 3 force components of contact c
    __device__ __forceinline__ V3<R> fric_coeff(int c, int which) const {
        const V3<R> n = ld3(PR.cn[c]), s0 = ld3(PR.cspan[c]), s1 = ld3(PR.cspan[c] + 3);
        if (which == 0) return n;
        const R a = (which >= 3) ? R(1) : R(-1), b = (which == 2 || which == 4) ? R(1) : R(-1);
        return PR.cmu[c] * n + a * s0 + b * s1;
    }
    // value of ineq row r of stage k at the QP iterate z (stage vector zk = [du; dx]), and bounds.
    // Uses the linearisation stored in the workspace.
    __device__ R row_value(int k, int r, int fam, const R* zk, const F* xk, const F* uk, R* lb, R* ub) const {
        const int nq = NQ(), nu = NU();
        if (fam == 0) {
            const R u = R(uk[r]);
            *lb = (r < nq ? PR.ulb[r] : PR.flb) - u;
            *ub = (r < nq ? PR.uub[r] : PR.fub) - u;
            return zk[r];
        }
        if (fam == 1) {
            const int i = r - NBOXU();
            const R x = R(xk[i]);
            *lb = PR.xlb[i] - x;
            *ub = PR.xub[i] - x;
            return zk[nu + i];
        }
        *lb = R(0);
        *ub = tinf<R>();
        if (fam == 2) {
            const int i = r - NBOXU() - NX(), c = i / 5;
            const V3<R> a = fric_coeff(c, i % 5);
            const F* f = uk + nq + 3 * c;
            const R* df = zk + nq + 3 * c;
            return a.x * (R(f[0]) + df[0]) + a.y * (R(f[1]) + df[1]) + a.z * (R(f[2]) + df[2]);
        }
        const int i = r - NBOXU() - NX() - NFRIC();
        const int ow = OBSW();
        const F* J = ws + oLJO() + (k * NOBS() + i) * ow;
        R v = R(ws[oLHO() + k * NOBS() + i]);
        for (int j = 0; j < ow; ++j) v += R(J[j]) * zk[nu + j];
        return v;
    }
    // a_r . d  for a stage direction d = [du; dx] (a direction: F arithmetic)
    __device__ F row_dot(int k, int r, int fam, const F* d) const {
        const int nq = NQ(), nu = NU();
        if (fam == 0) return d[r];
        if (fam == 1) return d[nu + r - NBOXU()];
        if (fam == 2) {
            const int i = r - NBOXU() - NX(), c = i / 5;
            const V3<R> a = fric_coeff(c, i % 5);
            const F* df = d + nq + 3 * c;
            return F(a.x) * df[0] + F(a.y) * df[1] + F(a.z) * df[2];
        }
        const int i = r - NBOXU() - NX() - NFRIC();
        const int ow = OBSW();
        const F* J = ws + oLJO() + (k * NOBS() + i) * ow;
        F v = 0;
        for (int j = 0; j < ow; ++j) v += J[j] * d[nu + j];
        return v;
    }
    // Slack/multiplier record of one inequality row: {t_lo, t_hi, lam_lo, lam_hi} (double) and the step
    // {dt_lo, dt_hi, dlam_lo, dlam_hi} (F), consecutive per row so that a warp reads the rows of a stage with
    // fully coalesced vector loads.
    struct alignas(16) QuadR {
        R v[4];
    };
    struct alignas(16) QuadF {
        F v[4];
    };
    __device__ __forceinline__ QuadR* side_tl(int k, int r) const { return wsr<QuadR>(oTL()) + (k * NROW() + r); }
    __device__ __forceinline__ QuadF* side_dd(int k, int r) const { return wsr<QuadF>(oDD()) + (k * NROW() + r); }
    // Staging of the side records: the records of the stage a pass visits NEXT are copied into shared memory
    // with cp.async while the current stage computes.  Writers always store to the workspace.
    static constexpr bool kStageTT = D::kStatic && D::nrow <= UB_STAGE_ROWS_MAX;
    // every specialised kernel prefetches the factor block (into the idle half of the stage-matrix buffer), the
    // force bundle and w_k of the stage a sweep visits next: they need no buffer of their own
    static constexpr bool kStageFB = D::kStatic;
    __device__ __forceinline__ const QuadR* recs_tl(int k) const {
        if constexpr (kStageTT) return reinterpret_cast<const QuadR*>(sTL);
        else return wsr<QuadR>(oTL()) + k * NROW();
    }
    __device__ __forceinline__ const QuadF* recs_dd(int k) const {
        if constexpr (kStageTT) return reinterpret_cast<const QuadF*>(sDD);
        else return wsr<QuadF>(oDD()) + k * NROW();
    }
    __device__ __forceinline__ void cp_async16(void* dst, const void* src) const {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    }
#if UB_TMA_STAGE
    // Variant (A/B experiment, profiles/r2_v7_tma_ab.txt): the 16-byte-granular blocks of a stage (factor block, side
    // records, force bundle) travel as ONE bulk copy each (cp.async.bulk, the non-tensor TMA path) issued by lane 0 and
    // complete on an mbarrier of this warp; the small vectors stay on cp.async.  A "group" is what the commit-group
    // mechanism calls one: group g completes on mbarrier g & 1 in phase (g >> 1) & 1; at most two groups are in flight
    // (every third commit is preceded by a wait that retires the oldest).
    mutable unsigned gcount = 0, gwaited = 0;
    uint32_t bar0 = 0;   // shared-memory address of the two mbarriers of this warp
    __device__ __forceinline__ void bar_init(void* bars) {
        bar0 = smem_u32(bars);
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        tsync();
    }
    // every lane fences its own generic-proxy stores of the previous pass against the async-proxy reads of this one
    __device__ __forceinline__ void pass_fence() const {
        asm volatile("fence.proxy.async;" ::: "memory");
        tsync();
    }
    __device__ __forceinline__ void cp_commit() const {
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (lane == 0) {
            unsigned long long st;
            asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(bar0 + 8 * (gcount & 1)) : "memory");
        }
        ++gcount;
        if (gcount - gwaited > 2) __trap();   // a barrier would be re-armed before its previous phase was retired
    }
    template <int NYOUNGER>
    __device__ __forceinline__ void cp_wait() const {
        asm volatile("cp.async.wait_group %0;" ::"n"(NYOUNGER) : "memory");
        while (gwaited + NYOUNGER < gcount) {
            const uint32_t bar = bar0 + 8 * (gwaited & 1), parity = (gwaited >> 1) & 1;
            uint32_t done = 0;
            int spins = 0;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                if (++spins > (1 << 20)) __trap();   // a lost completion must not hang the GPU
            } while (!done);
            ++gwaited;
        }
        tsync();
    }
    __device__ __forceinline__ void cp_async_bytes(void* dst, const void* src, int bytes) const {
        if (lane == 0) {
            const uint32_t bar = bar0 + 8 * (gcount & 1);
            asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
        }
    }
#else
    __device__ __forceinline__ void bar_init(void*) {}
    __device__ __forceinline__ void pass_fence() const {}
    __device__ __forceinline__ void cp_commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
    template <int NYOUNGER>
    __device__ __forceinline__ void cp_wait() const {
        asm volatile("cp.async.wait_group %0;" ::"n"(NYOUNGER) : "memory");
        tsync();
    }
    // bytes -> 16-byte chunks, both sides 16-byte aligned (every block of the layouts is)
    __device__ __forceinline__ void cp_async_bytes(void* dst, const void* src, int bytes) const {
        const char* s = reinterpret_cast<const char*>(src);
        char* d = reinterpret_cast<char*>(dst);
        for (int i = lane; i < bytes / 16; i += kTS) cp_async16(d + 16 * i, s + 16 * i);
    }
#endif
    // issue (no commit) the copy of the records of stage k; k outside [0, N] issues nothing
    __device__ __forceinline__ void tt_issue(int k, bool with_steps) const {
        if constexpr (kStageTT) {
            if (k < 0 || k > NN()) return;
            cp_async_bytes(sTL, wsr<QuadR>(oTL()) + k * D::nrow, D::nrow * 4 * int(sizeof(R)));
            if (with_steps) cp_async_bytes(sDD, wsr<QuadF>(oDD()) + k * D::nrow, D::nrow * 4 * int(sizeof(F)));
        } else if constexpr (kStageFB) {
            // stages too wide to stage in shared memory (cfg3: 164 rows, cfg5: 68): pull the records of the next stage
            // into L2 at least, so that the loads of the next stage do not wait for DRAM
            if (k < 0 || k > NN()) return;
            l2_prefetch(wsr<QuadR>(oTL()) + k * D::nrow, D::nrow * 4 * int(sizeof(R)));
            if (with_steps) l2_prefetch(wsr<QuadF>(oDD()) + k * D::nrow, D::nrow * 4 * int(sizeof(F)));
        }
    }
    __device__ __forceinline__ void l2_prefetch(const void* src, int bytes) const {
        const char* s = reinterpret_cast<const char*>(src);
        for (int i = lane * 128; i < bytes; i += kTS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(s + i));
    }
    // staged per-stage vectors (same schedule as the side records): the QP iterate z_k (or, in the corrector
    // pass, the stored predictor gradient), the linearisation point x_k, u_k, the position Jacobian and w_k
    __device__ __forceinline__ R* Zk(int k) const { return wsr<R>(oZ()) + k * NZ(); }
    __device__ __forceinline__ F* DZk(int k) const { return ws + oDZ() + k * NZ(); }
    __device__ __forceinline__ R* GPk(int k) const { return wsr<R>(oGP()) + k * NZ(); }
    __device__ __forceinline__ const R* st_z(int k) const { if constexpr (kStageTT) return sSmZ; else return Zk(k); }
    __device__ __forceinline__ const R* st_gp(int k) const { if constexpr (kStageTT) return sSmZ; else return GPk(k); }
    __device__ __forceinline__ const F* st_x(int k) const { if constexpr (kStageTT) return sSmX; else return X + k * NX(); }
    __device__ __forceinline__ const F* st_u(int k) const { if constexpr (kStageTT) return sSmU; else return U + k * NU(); }
    __device__ __forceinline__ const F* st_jp(int k) const { if constexpr (kStageTT) return sSmJ; else return ws + oLJP() + k * 3 * NQ(); }
    template <typename T>
    __device__ __forceinline__ void cp_async_elems(T* dst, const T* src, int n) const {
        for (int i = lane; i < n; i += kTS) {
            if constexpr (sizeof(T) == 4)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + i)), "l"(src + i) : "memory");
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst + i)), "l"(src + i) : "memory");
        }
    }
    // z_k (gp = false) or the predictor gradient of stage k (gp = true), x_k, u_k and optionally Jp_k
    __device__ __forceinline__ void sm_issue(int k, bool gp, bool jp) const {
        if constexpr (!kStageTT && kStageFB) {
            if (k < 0 || k > NN()) return;
            l2_prefetch(gp ? GPk(k) : Zk(k), NZ() * int(sizeof(R)));
        }
        if constexpr (kStageTT) {
            if (k < 0 || k > NN()) return;
            cp_async_elems(sSmZ, gp ? GPk(k) : Zk(k), NZ());
            if (!gp) {
                cp_async_elems(sSmX, X + k * NX(), NX());
                if (k < NN()) cp_async_elems(sSmU, U + k * NU(), NU());
            }
            if (jp) cp_async_elems(sSmJ, ws + oLJP() + k * 3 * NQ(), 3 * NQ());
        }
    }
    __device__ __forceinline__ void w_issue(int k) const {
        if constexpr (kStageFB) {
            if (k < 0 || k >= NN()) return;
            cp_async_elems(sSmW, ws + oWF() + k * NQ(), NQ());
        }
    }
    // Newton data of one side: the coefficient that multiplies sgn*a in the stage gradient;
    // d = signed distance to the bound at the current iterate
    // Only the slack residual rd (operands O(1), result -> 0) and the accumulation of -lam need double: the
    // remaining term is small near the solution and is formed in F.
    __device__ __forceinline__ R side_coef(R t, R lam, R d, R eps, F target, F corr) const {
        const R rd = d + eps * lam - t;
        const F tf = F(t), lf = F(lam);
        const F rc = tf * lf - target + corr;
        return -lam + R(fdiv(rc + lf * F(rd), tf + F(eps) * lf));
    }
    __device__ __forceinline__ int neq_of(int k) const { return k < NN() ? NEQ() : NTERM(); }

    // -------------------------------------------------------- linearisation
    // Df: d g / d f, constant in x (compute_object_wrenches, contact_constraints.h:106-157): dense [neq][nfc] and
    // per contact and side [c][side][6][nf] (side 0 = the body the force acts on as object 2, side 1 = object 1)
    __device__ void build_Df() {
        const R scale = rsqrt(R(6 * NB()));
        F* Df = ws + oDF();
        F* Dc = sDFC;
        const int nf = NF(), nfc = NFC();
        for (int idx = lane; idx < NEQ() * nfc; idx += kTS) Df[idx] = F(0);
        for (int idx = lane; idx < NC() * 12 * nf; idx += kTS) Dc[idx] = F(0);
        tsync();
        for (int j = lane; j < nfc; j += kTS) {
            const int c = j / nf, comp = j % nf;
            V3<R> e;
            if (nf == 1) e = ld3(PR.cn[c]);
            else e = V3<R>(comp == 0 ? R(1) : R(0), comp == 1 ? R(1) : R(0), comp == 2 ? R(1) : R(0));
            const int b1 = P.cb1[c], b2 = P.cb2[c];
            if (b1 >= 0) {
                const BodyP<R> Bd = load_body<R>(body + b1 * UB_BODY_PARAMS);
                const V3<R> tq = cross(ld3(PR.cr1[c]) - Bd.com, e);
                const R s = -scale / Bd.m;
                const R col[6] = {s * e.x, s * e.y, s * e.z, s * tq.x, s * tq.y, s * tq.z};
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    Df[(6 * b1 + i) * nfc + j] = F(col[i]);
                    Dc[((2 * c + 1) * 6 + i) * nf + comp] = F(col[i]);
                }
            }
            {
                const BodyP<R> Bd = load_body<R>(body + b2 * UB_BODY_PARAMS);
                const V3<R> tq = cross(ld3(PR.cr2[c]) - Bd.com, R(-1) * e);
                const R s = -scale / Bd.m;
                const R col[6] = {-s * e.x, -s * e.y, -s * e.z, s * tq.x, s * tq.y, s * tq.z};
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    Df[(6 * b2 + i) * nfc + j] = F(col[i]);
                    Dc[((2 * c) * 6 + i) * nf + comp] = F(col[i]);
                }
            }
        }
        tsync();
    }
    // (Df v)[row] for a force-space vector v (shared memory): the contacts of the row's body from the per-body lists
    template <typename T>
    __device__ __forceinline__ R eq_force_dot(int row, const T* v) const {
        const int b = row / 6, rr = row - 6 * b, nf = NF();
        const F* Dc = sDFC;
        R acc = 0;
#pragma unroll 1
        for (int e = P.bc_start[b]; e < P.bc_start[b + 1]; ++e) {
            const int cs = P.bc_list[e], c = cs >> 1;
            const F* d = Dc + (cs * 6 + rr) * nf;
#pragma unroll 1
            for (int i = 0; i < nf; ++i) acc += R(d[i]) * R(v[c * nf + i]);
        }
        return acc;
    }

    // Linearise every knot around (X, U): lane j carries d/dx_j.  Computed in double (the QP solution is far more
    // sensitive to the linearisation than to anything else: 1e-7 relative noise in these blocks moves the hard-constraint
    // configurations by 1e-3 of the limit range), stored in F except the values the residuals are built from.
    // Writes LG [k][neq] (g value incl. Df f), LC [k][neq][nx], LR [k][3], LJP [k][3][nq], LHO [k][nobs],
    // LJO [k][nobs][obsw], GAP [k][nx].
    __device__ void linearize() {
        const int nq = NQ(), nx = NX(), nu = NU(), N = NN();
        const R scale = rsqrt(R(6 * max(NB(), 1)));
        R sph[3 * UB_MAX_SPHERES], dsph[3 * UB_MAX_SPHERES];
        // dynamics gap b_k = A x_k + B u_k - x_{k+1}  (exact triple integrator, system_dynamics.h:15-26)
        auto write_gap = [&](int k) {
            if (k < N && lane < nq) {
                const F* x = X + k * nx;
                const R dt = PR.dt;
                const F* xn = X + (k + 1) * nx;
                const R q = R(x[lane]), v = R(x[nq + lane]), a = R(x[2 * nq + lane]), j = R(U[k * nu + lane]);
                R* gap = wsr<R>(oGAP()) + k * nx;
                gap[lane] = q + dt * v + R(0.5) * dt * dt * a + dt * dt * dt / R(6) * j - R(xn[lane]);
                gap[nq + lane] = v + dt * a + R(0.5) * dt * dt * j - R(xn[nq + lane]);
                gap[2 * nq + lane] = a + dt * j - R(xn[2 * nq + lane]);
            }
        };
        // A knot whose (x, u) equals the previous knot's bit for bit has the same linearisation: copied, not recomputed.
        // (Every knot of a cold start — DefaultInitializer: state held, zero input — is such a knot.)  Rows that depend
        // on the knot index itself (orientation / box targets, dynamic obstacles, projectile) switch the shortcut off.
        const bool memo_ok = !(ORI() || IALIGN() || IACON() || EEBOX() || NXO() > 0 || NPROJ() > 0);
        for (int k = 0; k <= N; ++k) {
            const F* x = X + k * nx;
            if (memo_ok && k > 0) {
                bool eq = true;
                const F* xp = x - nx;
                for (int i = lane; i < nx; i += kTS) eq = eq && (x[i] == xp[i]);
                if (k < N) {
                    const F* u = U + k * nu;
                    for (int i = lane; i < nu; i += kTS) eq = eq && (u[i] == u[i - nu]);
                }
                if (__all_sync(FULL, eq)) {
                    tsync();   // the blocks of knot k - 1 were written by other lanes
                    const int no = NOBS(), ne = NEQ();
                    if (lane < 3) wsr<R>(oLR())[3 * k + lane] = wsr<R>(oLR())[3 * (k - 1) + lane];
                    for (int i = lane; i < 3 * nq; i += kTS) ws[oLJP() + k * 3 * nq + i] = ws[oLJP() + (k - 1) * 3 * nq + i];
                    for (int i = lane; i < no; i += kTS) ws[oLHO() + k * no + i] = ws[oLHO() + (k - 1) * no + i];
                    for (int i = lane; i < no * OBSW(); i += kTS) ws[oLJO() + k * no * OBSW() + i] = ws[oLJO() + (k - 1) * no * OBSW() + i];
                    if (k < N) {
                        for (int i = lane; i < ne * nx; i += kTS) ws[oLC() + k * ne * nx + i] = ws[oLC() + (k - 1) * ne * nx + i];
                        for (int i = lane; i < ne; i += kTS) wsr<R>(oLG())[k * ne + i] = wsr<R>(oLG())[(k - 1) * ne + i];
                    }
                    write_gap(k);
                    tsync();
                    continue;
                }
            }
            Kin<R> Kn;
            KinTan<R> Dt;
            // sin / cos of the joint angles once per knot (lane i: joint i), shared through the scratch
            tsync();
            if (lane < nq) {
                R sn, cs;
                sincos(R(x[lane]), &sn, &cs);
                sScr[2 * lane] = sn;
                sScr[2 * lane + 1] = cs;
            }
            tsync();
            forward_kinematics<R, true, F>(PR, x, lane, Kn, Dt, SPHERES() ? sph : nullptr, dsph, sScr);
            if (NXO() > 0) place_dynamic_spheres<R>(k, R(0), sph);
            if (lane == 0) {
                R* lr = wsr<R>(oLR()) + 3 * k;
                lr[0] = Kn.r.x;
                lr[1] = Kn.r.y;
                lr[2] = Kn.r.z;
            }
            if (lane < nq) {
                F* Jp = ws + oLJP() + k * 3 * nq;
                Jp[lane] = F(Dt.r.x);
                Jp[nq + lane] = F(Dt.r.y);
                Jp[2 * nq + lane] = F(Dt.r.z);
            }
            if (ORI() && k < N) {
                R qr[4];
                for (int c = 0; c < 4; ++c) qr[c] = R(ws[oTGQ() + 4 * k + c]);
                V3<R> de;
                const V3<R> eo = orientation_error<R, true>(Kn.C, Dt.th, qr, &de);
                if (lane == 0) {
                    R* lo = wsr<R>(oLRO()) + 3 * k;
                    lo[0] = eo.x;
                    lo[1] = eo.y;
                    lo[2] = eo.z;
                }
                if (lane < nq) {
                    F* Jq = ws + oLJQ() + k * 3 * nq;
                    Jq[lane] = F(de.x);
                    Jq[nq + lane] = F(de.y);
                    Jq[2 * nq + lane] = F(de.z);
                }
            }
            if (k < N && NEQ() > 0) {
                for (int b = 0; b < NB(); ++b) {
                    const BodyP<R> Bd = load_body<R>(body + b * UB_BODY_PARAMS);
                    R g6[6], dg6[6];
                    object_dynamics_state_part<R, true>(PR, Bd, Kn, Dt, scale, g6, dg6);
                    if (lane < nx) {
                        F* Rw = ws + oLC() + (k * NEQ() + 6 * b) * nx + lane;
#pragma unroll
                        for (int i = 0; i < 6; ++i) Rw[i * nx] = F(dg6[i]);
                    }
                    if (lane < 6) {
                        // g = state part + Df f
                        R gv = g6[0];
#pragma unroll
                        for (int i = 1; i < 6; ++i) gv = (lane == i) ? g6[i] : gv;
                        gv += eq_force_dot(6 * b + lane, U + k * nu + nq);
                        wsr<R>(oLG())[k * NEQ() + 6 * b + lane] = gv;
                    }
                }
            }
            if (IALIGN() && k < N) {
                // e = S C_we' (a - g) / |g| and its Jacobian (inertial_alignment.cpp:151-163)
                R e2[2], de2[2];
                inertial_alignment_error<R, true>(PR, Kn, Dt, e2, de2);
                if (lane == 0) {
                    ws[oLIA() + 2 * k] = F(e2[0]);
                    ws[oLIA() + 2 * k + 1] = F(e2[1]);
                }
                if (lane < nx) {
                    ws[oLJA() + (2 * k) * nx + lane] = F(de2[0]);
                    ws[oLJA() + (2 * k + 1) * nx + lane] = F(de2[1]);
                }
            }
            if (NOBS() > 0) {
                for (int i = 0; i < NPAIRS(); ++i) {
                    const int a = P.pa[i], bb = P.pb[i];
                    V3<R> dir;   // d h / d c_a = - d h / d c_b
                    const R sep = pair_separation(PR, a, bb, sph, &dir);
                    const V3<R> dd(dsph[3 * a] - dsph[3 * bb], dsph[3 * a + 1] - dsph[3 * bb + 1],
                                   dsph[3 * a + 2] - dsph[3 * bb + 2]);
                    R shift = R(0);
                    if (NXO() > 0) {
                        // known Newton step of the obstacle positions: h + (dh/dc_a) dp_a + (dh/dc_b) dp_b
                        for (int side = 0; side < 2; ++side) {
                            const int s = side == 0 ? a : bb;
                            if (P.slink[s] > -2) continue;
                            const F* dp = ws + oDXO() + (k * P.ndyn + (-2 - P.slink[s])) * 9;
                            const R proj = dir.x * R(dp[0]) + dir.y * R(dp[1]) + dir.z * R(dp[2]);
                            shift += side == 0 ? proj : -proj;
                        }
                    }
                    if (lane == 0) ws[oLHO() + k * NOBS() + i] = F(sep - PR.dmin + shift);
                    if (lane < OBSW()) ws[oLJO() + (k * NOBS() + i) * OBSW() + lane] = lane < nq ? F(dot(dir, dd)) : F(0);
                }
                if (EEBOX()) {
                    // rows npairs..+2: r_d + upper - r >= 0; rows npairs+3..+5: r - r_d - lower >= 0
                    // (end_effector_box_constraint.h:46-76)
                    const F* tg = target + 3 * k;
                    for (int c = 0; c < 3; ++c) {
                        const int iu = NPAIRS() + c, il = NPAIRS() + 3 + c;
                        if (lane == 0) {
                            ws[oLHO() + k * NOBS() + iu] = F(R(tg[c]) + PR.eb_hi[c] - Kn.r[c]);
                            ws[oLHO() + k * NOBS() + il] = F(Kn.r[c] - R(tg[c]) - PR.eb_lo[c]);
                        }
                        if (lane < OBSW()) {
                            ws[oLJO() + (k * NOBS() + iu) * OBSW() + lane] = lane < nq ? F(-Dt.r[c]) : F(0);
                            ws[oLJO() + (k * NOBS() + il) * OBSW() + lane] = lane < nq ? F(Dt.r[c]) : F(0);
                        }
                    }
                }
                if (IACON()) {
                    // five inertial-alignment rows (inertial_alignment.cpp:7-53) behind the box rows, dense over x
                    R h5[5], dh5[5];
                    inertial_alignment_rows<R, true>(PR, Kn, Dt, h5, dh5);
                    const int i0 = NPAIRS() + (EEBOX() ? 6 : 0);
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        if (lane == 0) ws[oLHO() + k * NOBS() + i0 + r] = F(h5[r]);
                        if (lane < nx) ws[oLJO() + (k * NOBS() + i0 + r) * OBSW() + lane] = F(dh5[r]);
                    }
                }
                if (NPROJ() > 0) {
                    // projectile-path rows (projectile_path_constraint.h:108-146) close the family: dense over q with
                    // the time of closest approach held fixed; the obstacle-state block - w s n' [I, t I, t^2/2 I]
                    // meets the known Newton step of the (last) obstacle and lands in the constant
                    const int i0 = NPAIRS() + (EEBOX() ? 6 : 0) + (IACON() ? 5 : 0);
                    const int o = (k * P.ndyn + P.ndyn - 1) * 9;
                    R xo[9], dxo[9];
                    for (int c = 0; c < 9; ++c) {
                        xo[c] = R(ws[oXO() + o + c]);
                        dxo[c] = R(ws[oDXO() + o + c]);
                    }
                    for (int i = 0; i < NPROJ(); ++i) {
                        const int a = P.proj_sph[i];
                        V3<R> n;
                        R tc;
                        const R h = projectile_row(PR, i, ld3(sph + 3 * a), xo, &n, &tc);
                        const R w = PR.proj_scale / PR.proj_d[i] * PR.proj_s;
                        const V3<R> dstep = ld3(dxo) + tc * ld3(dxo + 3) + (R(0.5) * tc * tc) * ld3(dxo + 6);
                        if (lane == 0) ws[oLHO() + k * NOBS() + i0 + i] = F(h - w * dot(n, dstep));
                        if (lane < OBSW())
                            ws[oLJO() + (k * NOBS() + i0 + i) * OBSW() + lane] = lane < nq ? F(w * dot(n, ld3(dsph + 3 * a))) : F(0);
                    }
                }
            }
            write_gap(k);
        }
        tsync();
    }

    // --------------------------------------------------- performance index
    // Lane k evaluates knot k (values only).  Mirrors orc::performance().
    // `ao`: step along the (known) obstacle-state direction, 0 for the current iterate
    __device__ Perf<F> performance(const F* Xt, const F* Ut, F ao = F(0)) const {
        const int nq = NQ(), nx = NX(), nu = NU(), N = NN();
        const F dt = C.dt;
        const F scale = rsqrt(F(6 * max(NB(), 1)));
        F cost = 0, dyn = 0, eq = 0, ineq = 0, max_eq = 0, min_margin = tinf<F>();
        F sph[3 * UB_MAX_SPHERES];
        for (int k = lane; k <= N; k += kTS) {
            const F* x = Xt + k * nx;
            Kin<F> Kn;
            KinTan<F> Dn;
            forward_kinematics<F, false>(P, x, -1, Kn, Dn, SPHERES() ? sph : nullptr, nullptr);
            if (NXO() > 0) {
                place_dynamic_spheres<F>(k, ao, sph);
                if (k < N)   // dynamics defect of the obstacle states: (1 - ao) x the defect of the iterate
                    for (int j = 0; j < P.ndyn; ++j) {
                        const F* o0 = ws + oXO() + (k * P.ndyn + j) * 9;
                        const F* o1 = ws + oXO() + ((k + 1) * P.ndyn + j) * 9;
                        for (int c = 0; c < 3; ++c) {
                            const F g0 = o0[c] + dt * o0[3 + c] + F(0.5) * dt * dt * o0[6 + c] - o1[c];
                            const F g1 = o0[3 + c] + dt * o0[6 + c] - o1[3 + c];
                            const F g2 = o0[6 + c] - o1[6 + c];
                            dyn += dt * (F(1) - ao) * (F(1) - ao) * (g0 * g0 + g1 * g1 + g2 * g2);
                        }
                    }
            }
            const F* rd = target + 3 * k;
            if (k == N) {
                for (int i = 0; i < 3; ++i) {
                    const F e = rd[i] - Kn.r[i];
                    eq += e * e;
                    max_eq = max(max_eq, fabs(e));
                }
                for (int i = nq; i < nx; ++i) {
                    eq += x[i] * x[i];
                    max_eq = max(max_eq, fabs(x[i]));
                }
            }
            if (k >= 1)
                for (int i = 0; i < nx; ++i) {
                    const F lo = x[i] - P.xlb[i], hi = P.xub[i] - x[i];
                    const F a = min(F(0), lo), b = min(F(0), hi);
                    ineq += dt * (a * a + b * b);
                    min_margin = min(min_margin, min(lo, hi));
                }
            if (k == N) continue;
            const F* u = Ut + k * nu;
            F c = 0;
            for (int i = 0; i < nx; ++i) {
                const F e = x[i] - P.xd[i];
                c += F(0.5) * P.Qd[i] * e * e;
            }
            for (int i = 0; i < nq; ++i) c += F(0.5) * P.Rd[i] * u[i] * u[i];
            for (int i = 0; i < NFC(); ++i) c += F(0.5) * C.fw * u[nq + i] * u[nq + i];
            for (int i = 0; i < 3; ++i) {
                const F e = Kn.r[i] - rd[i];
                c += F(0.5) * P.Wd[i] * e * e;
            }
            if (IALIGN()) {
                F e2[2];
                inertial_alignment_error<F, false>(P, Kn, Dn, e2, nullptr);
                c += F(0.5) * P.ia_w * (e2[0] * e2[0] + e2[1] * e2[1]);
            }
            if (ORI()) {
                const V3<F> eo = orientation_error<F, false>(Kn.C, V3<F>(), ws + oTGQ() + 4 * k, nullptr);
                c += F(0.5) * (P.Wo[0] * eo.x * eo.x + P.Wo[1] * eo.y * eo.y + P.Wo[2] * eo.z * eo.z);
            }
            cost += dt * c;
            const F* xn = Xt + (k + 1) * nx;
            for (int i = 0; i < nq; ++i) {
                const F q = x[i], v = x[nq + i], a = x[2 * nq + i], j = u[i];
                const F g0 = q + dt * v + F(0.5) * dt * dt * a + dt * dt * dt / F(6) * j - xn[i];
                const F g1 = v + dt * a + F(0.5) * dt * dt * j - xn[nq + i];
                const F g2 = a + dt * j - xn[2 * nq + i];
                dyn += dt * (g0 * g0 + g1 * g1 + g2 * g2);
            }
            const int nbox = NBOXU();
            for (int i = 0; i < nbox; ++i) {
                const F lo = u[i] - (i < nq ? P.ulb[i] : C.flb), hi = (i < nq ? P.uub[i] : C.fub) - u[i];
                const F a = min(F(0), lo), b = min(F(0), hi);
                ineq += dt * (a * a + b * b);
                min_margin = min(min_margin, min(lo, hi));
            }
            for (int b = 0; b < (NEQ() > 0 ? NB() : 0); ++b) {
                const BodyP<F> Bd = load_body<F>(body + b * UB_BODY_PARAMS);
                F g6[6];
                object_dynamics_state_part<F, false>(P, Bd, Kn, Dn, scale, g6, nullptr);
                for (int i = 0; i < 6; ++i) {
                    const F* Dfr = ws + oDF() + (6 * b + i) * NFC();
                    F gv = g6[i];
                    for (int j = 0; j < NFC(); ++j) gv += Dfr[j] * u[nq + j];
                    eq += dt * gv * gv;
                    max_eq = max(max_eq, fabs(gv));
                }
            }
            for (int i = 0; i < NFRIC(); ++i) {
                const int cidx = i / 5, which = i % 5;
                const V3<F> n = ld3(P.cn[cidx]), s0 = ld3(P.cspan[cidx]), s1 = ld3(P.cspan[cidx] + 3);
                V3<F> a = n;
                if (which > 0) {
                    const F sa = (which >= 3) ? F(1) : F(-1), sb = (which == 2 || which == 4) ? F(1) : F(-1);
                    a = P.cmu[cidx] * n + sa * s0 + sb * s1;
                }
                const F* f = u + nq + 3 * cidx;
                const F h = a.x * f[0] + a.y * f[1] + a.z * f[2];
                const F m = min(F(0), h);
                ineq += dt * m * m;
                min_margin = min(min_margin, h);
            }
            if (k >= 1)
            {
                for (int i = 0; i < NPAIRS(); ++i) {
                    V3<F> dir;
                    const F h = pair_separation(P, P.pa[i], P.pb[i], sph, &dir) - C.dmin;
                    const F m = min(F(0), h);
                    ineq += dt * m * m;
                    min_margin = min(min_margin, h);
                }
                if (EEBOX())
                    for (int c = 0; c < 3; ++c) {
                        const F hu = rd[c] + P.eb_hi[c] - Kn.r[c], hl = Kn.r[c] - rd[c] - P.eb_lo[c];
                        const F mu_ = min(F(0), hu), ml = min(F(0), hl);
                        ineq += dt * (mu_ * mu_ + ml * ml);
                        min_margin = min(min_margin, min(hu, hl));
                    }
                if (IACON()) {
                    F h5[5];
                    inertial_alignment_rows<F, false>(P, Kn, Dn, h5, nullptr);
                    for (int r = 0; r < 5; ++r) {
                        const F m = min(F(0), h5[r]);
                        ineq += dt * m * m;
                        min_margin = min(min_margin, h5[r]);
                    }
                }
                if (NPROJ() > 0) {
                    const int o = (k * P.ndyn + P.ndyn - 1) * 9;
                    F xo[9];
                    for (int c = 0; c < 9; ++c) xo[c] = ws[oXO() + o + c] + (ao != F(0) ? ao * ws[oDXO() + o + c] : F(0));
                    for (int i = 0; i < NPROJ(); ++i) {
                        V3<F> n;
                        F tc;
                        const F h = projectile_row(P, i, ld3(sph + 3 * P.proj_sph[i]), xo, &n, &tc);
                        const F m = min(F(0), h);
                        ineq += dt * m * m;
                        min_margin = min(min_margin, h);
                    }
                }
            }
        }
        Perf<F> pf;
        pf.cost = tsum(cost);
        pf.dyn = tsum(dyn);
        pf.eq = tsum(eq);
        pf.ineq = tsum(ineq);
        pf.max_eq = tmax(max_eq);
        pf.min_margin = tmin(min_margin);
        return pf;
    }

    // ------------------------------------------------------------ equality rows
    // Stage k < N: object-dynamics rows [0 | Df | C], eliminated with the forces (force block below).
    // Stage N: terminal rows [r_d - r; v; a] = 0 (stationary_desired_position_constraint.h:43-74): three dense rows
    // over q (-Jp) and unit rows, kept as weighted terms of the state block.
    __device__ __forceinline__ R* rho_eq(int k) const { return k < NN() ? wsr<R>(oRHOE()) + k * NEQ() : wsr<R>(oRHOT()); }
    __device__ __forceinline__ R* y_eq(int k) const { return k < NN() ? wsr<R>(oYE()) + k * NEQ() : wsr<R>(oYT()); }
    // the C rows of stage k into shared memory (dense [neq][nx]): synchronously, or issued as cp.async (no commit)
    __device__ __forceinline__ void load_C(int k) {
        const int nx = NX(), ne = NEQ();
        const F* __restrict__ src = ws + oLC() + k * ne * nx;
        for (int idx = lane; idx < ne * nx; idx += kTS) cC[idx] = src[idx];
        tsync();
    }
    __device__ __forceinline__ void c_issue(int k) const {
        if constexpr (!kStageTT && kStageFB) {
            if (NEQ() == 0 || k < 0 || k >= NN()) return;
            l2_prefetch(ws + oLC() + k * D::neq * D::nx, D::neq * D::nx * int(sizeof(F)));
        }
        if constexpr (kStageTT) {
            if (NEQ() == 0 || k < 0 || k >= NN()) return;
            cp_async_elems(sCst, ws + oLC() + k * D::neq * D::nx, D::neq * D::nx);   // stage blocks are only 4-byte aligned
        }
    }
    // 1 / rho of equality row i of stage k < N (soft rows: the slack weight, no load)
    __device__ __forceinline__ R rho_inv_stage(int k, int i) const {
        if (C.soft_poly) return PR.invZ;
        const R rho = rho_eq(k)[i];
        return rho > R(0) ? rinv(rho) : R(1e30);
    }
    // value of equality row i of stage k < N at the QP iterate zk (C row from shared memory)
    __device__ __forceinline__ R eq_value_stage(int k, int i, const R* zk) const {
        const int nx = NX(), nu = NU(), nq = NQ();
        R v = wsr<R>(oLG())[k * NEQ() + i];
        const F* c = cC + i * nx;
        for (int j = 0; j < nx; ++j) v += R(c[j]) * zk[nu + j];
        return v + eq_force_dot(i, zk + nq);
    }
    // values of all equality rows of stage k < N at zk -> out (shared memory).  Few rows (one body): four lanes share
    // a row's 3 nq-long dot product, which cuts the dependent chain to a quarter.
    __device__ __forceinline__ void eq_values_stage(int k, const R* zk, R* out) const {
        const int nx = NX(), nu = NU(), nq = NQ(), ne = NEQ();
        if (ne <= 8) {
            const int i = lane >> 2, part = lane & 3;
            R v = R(0);
            if (i < ne) {
                const F* c = cC + i * nx;
                for (int j = part; j < nx; j += 4) v += R(c[j]) * zk[nu + j];
                if (part == 0) v += wsr<R>(oLG())[k * ne + i] + eq_force_dot(i, zk + nq);
            }
            v += __shfl_xor_sync(FULL, v, 1);
            v += __shfl_xor_sync(FULL, v, 2);
            if (i < ne && part == 0) out[i] = v;
        } else {
            for (int i = lane; i < ne; i += kTS) out[i] = eq_value_stage(k, i, zk);
        }
    }
    // value of terminal row i at the QP iterate zk
    __device__ __forceinline__ R eq_value_term(int i, const R* zk) const {
        const int nq = NQ(), nu = NU(), N = NN();
        if (i >= 3) return zk[nu + nq + (i - 3)] + R(X[N * NX() + nq + (i - 3)]);
        R v = R(target[3 * N + i]) - wsr<R>(oLR())[3 * N + i];
        const F* Jp = ws + oLJP() + (N * 3 + i) * nq;
        for (int j = 0; j < nq; ++j) v -= R(Jp[j]) * zk[nu + j];
        return v;
    }
    // Set the weights of the equality rows (soft: Z; hard: rho_hard on the unit-normalised row) and reset the
    // multipliers.
    __device__ void init_eq_weights() {
        const int nx = NX(), nq = NQ(), N = NN();
        for (int idx = lane; idx < N * NEQ(); idx += kTS) {
            const int k = idx / NEQ(), i = idx % NEQ();
            R rho = PR.Z;
            if (!C.soft_poly) {
                R n2 = 0;
                const F* c = ws + oLC() + (k * NEQ() + i) * nx;
                for (int j = 0; j < nx; ++j) n2 += R(c[j]) * R(c[j]);
                const F* df = ws + oDF() + i * NFC();
                for (int j = 0; j < NFC(); ++j) n2 += R(df[j]) * R(df[j]);
                rho = n2 > R(0) ? PR.rho_hard / n2 : R(0);
            }
            wsr<R>(oRHOE())[idx] = rho;
            wsr<R>(oYE())[idx] = R(0);
        }
        for (int i = lane; i < NTERM(); i += kTS) {
            R rho = PR.Z;
            if (!C.soft_poly) {
                R n2 = R(1);
                if (i < 3) {
                    n2 = 0;
                    const F* Jp = ws + oLJP() + (N * 3 + i) * nq;
                    for (int j = 0; j < nq; ++j) n2 += R(Jp[j]) * R(Jp[j]);
                }
                rho = n2 > R(0) ? PR.rho_hard / n2 : R(0);
            }
            wsr<R>(oRHOT())[i] = rho;
            wsr<R>(oYT())[i] = R(0);
        }
        tsync();
    }

    // M (lower triangle, ld = LDM()) = [B A]' Pn [B A] with the block structure A = A3 (x) I, B = B3 (x) I of the
    // exact triple-integrator discretisation: every (I >= J) block entry of the reduced stage matrix [j q v a] is
    // ASSIGNED (saves the separate zero fill of the buffer).  One pass over the nq x nq entry positions (rolled: the
    // body stays in the instruction cache): the nine P_ab(ii, jj) are loaded once and feed all ten block pairs; zero
    // entries of T3 drop out at compile time (52 products per position).
    __device__ void assign_dynamics_hessian() {
        constexpr int nq = D::nq, nx = D::nx, ld = LDM();
        const F dt = C.dt;
        // T3[a][I]: column 0 = B3, columns 1..3 = A3 (block order of the stage vector: jerk, q, v, a)
        const F T3[3][4] = {{dt * dt * dt / F(6), F(1), dt, F(0.5) * dt * dt},
                            {F(0.5) * dt * dt, F(0), F(1), dt},
                            {dt, F(0), F(0), F(1)}};
#pragma unroll 1
        for (int e = lane; e < nq * nq; e += kTS) {
            const int ii = e / nq, jj = e % nq;
            F pab[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) pab[a][b] = sP[(a * nq + ii) * nx + b * nq + jj];
#pragma unroll
            for (int I = 0; I < 4; ++I) {
#pragma unroll
                for (int J = 0; J <= I; ++J) {
                    F acc = F(0);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        if ((I == 1 && a != 0) || (I == 2 && a == 2)) continue;
#pragma unroll
                        for (int b = 0; b < 3; ++b) {
                            if ((J == 1 && b != 0) || (J == 2 && b == 2)) continue;
                            acc += T3[a][I] * T3[b][J] * pab[a][b];
                        }
                    }
                    sM[(I * nq + ii) * ld + J * nq + jj] = acc;   // diagonal blocks also fill (unused) upper entries
                }
            }
        }
        tsync();
    }
    // vec (reduced layout [j; x]) += [B A]' pv
    __device__ void add_dynamics_gradient(F* vec) const {
        constexpr int nq = D::nq;
        const F dt = C.dt;
        if (lane < nq) {
            const F p0 = sPv[lane], p1 = sPv[nq + lane], p2 = sPv[2 * nq + lane];
            vec[lane] += dt * dt * dt / F(6) * p0 + F(0.5) * dt * dt * p1 + dt * p2;
            vec[nq + lane] += p0;
            vec[2 * nq + lane] += dt * p0 + p1;
            vec[3 * nq + lane] += F(0.5) * dt * dt * p0 + dt * p1 + p2;
        }
        tsync();
    }

    // barrier weight lam / (t + eps lam) of both sides of a box row / the one side of a polytopic row
    // (matrix entries: F arithmetic; the force block wants them in double: *_r)
    __device__ __forceinline__ F box_weight(const QuadR& q, R eps) const {
        const F e = F(eps), l0 = F(q.v[2]), l1 = F(q.v[3]);
        return fdiv(l0, F(q.v[0]) + e * l0) + fdiv(l1, F(q.v[1]) + e * l1);
    }
    __device__ __forceinline__ F one_weight(const QuadR& q, R eps) const {
        const F l0 = F(q.v[2]);
        return fdiv(l0, F(q.v[0]) + F(eps) * l0);
    }
    __device__ __forceinline__ R box_weight_r(const QuadR& q, R eps) const {
        return q.v[2] * rinv(q.v[0] + eps * q.v[2]) + q.v[3] * rinv(q.v[1] + eps * q.v[3]);
    }
    __device__ __forceinline__ R one_weight_r(const QuadR& q, R eps) const { return q.v[2] * rinv(q.v[0] + eps * q.v[2]); }

    // Build the REDUCED Newton matrix of stage k in sM (lower triangle, [j; x]): cost Hessian + barrier terms of the
    // jerk / state / obstacle rows (+ the weighted terminal rows at k = N).  The force block and the object-dynamics
    // rows enter through force_block_factor().  `initialised`: the dynamics term has already been ASSIGNED.
    __device__ void build_stage_matrix(int k, bool initialised) {
        constexpr int nq = D::nq, nx = D::nx, nr = D::nr, ld = LDM();
        const F dt = C.dt;
        if (!initialised) {
            for (int idx = lane; idx < nr * ld; idx += kTS) sM[idx] = F(0);
            tsync();
        }
        const int nbu = NBOXU();
        // diagonal: cost weights + box barriers (one entry per lane and round: no conflicts)
        for (int i = lane; i < nr; i += kTS) {
            F d = F(0);
            if (k < NN()) d = i < nq ? dt * P.Rd[i] + C.reg_input : dt * P.Qd[i - nq];
            const int r = i < nq ? i : nbu + (i - nq);          // the box row of this variable
            const int fam = i < nq ? 0 : 1;
            if (row_valid(k, fam)) d += box_weight(recs_tl(k)[r], row_eps(fam));
            sM[i * ld + i] += d;
        }
        tsync();
        if (k < NN()) {
            // Gauss-Newton Hessian of the end-effector cost (end_effector_cost.h:48-84), scaled by dt
            const F* Jp = st_jp(k);
            for (int idx = lane; idx < nq * nq; idx += kTS) {
                const int a = idx / nq, b = idx % nq;
                if (b > a) continue;
                F acc = 0;
                for (int c = 0; c < 3; ++c) acc += C.Wd[c] * Jp[c * nq + a] * Jp[c * nq + b];
                if (ORI()) {
                    const F* Jq = ws + oLJQ() + k * 3 * nq;
                    for (int c = 0; c < 3; ++c) acc += P.Wo[c] * Jq[c * nq + a] * Jq[c * nq + b];
                }
                sM[(nq + a) * ld + nq + b] += dt * acc;
            }
            tsync();
            if (IALIGN()) {   // Gauss-Newton Hessian w Je' Je of the inertial-alignment cost (dense over x)
                const F* Ja = ws + oLJA() + 2 * k * nx;
                const F wa = dt * P.ia_w;
                for (int idx = lane; idx < nx * nx; idx += kTS) {
                    const int a = idx / nx, b = idx % nx;
                    if (b > a) continue;
                    sM[(nq + a) * ld + nq + b] += wa * (Ja[a] * Ja[b] + Ja[nx + a] * Ja[nx + b]);
                }
                tsync();
            }
        } else {
            // terminal rows: rho a a' (three dense rows over q, unit rows on v and a)
            const R* rho = rho_eq(k);
            const F* Jp = ws + oLJP() + k * 3 * nq;
            for (int idx = lane; idx < nq * nq; idx += kTS) {
                const int a = idx / nq, b = idx % nq;
                if (b > a) continue;
                F acc = 0;
                for (int c = 0; c < 3; ++c) acc += F(rho[c]) * Jp[c * nq + a] * Jp[c * nq + b];
                sM[(nq + a) * ld + nq + b] += acc;
            }
            tsync();
            for (int i = 3 + lane; i < NTERM(); i += kTS) {
                const int m = nq + nq + (i - 3);
                sM[m * ld + m] += F(rho[i]);
            }
            tsync();
        }
        if (NOBS() > 0 && k >= 1 && k < NN()) {
            const R eps = row_eps(3);
            const int nbx = nbu + nx;
            F* wrow = reinterpret_cast<F*>(sScr);  // barrier weights of the obstacle rows
            for (int i = lane; i < NOBS(); i += kTS) wrow[i] = one_weight(recs_tl(k)[nbx + NFRIC() + i], eps);
            tsync();
            const int ow = OBSW();
            for (int idx = lane; idx < ow * ow; idx += kTS) {
                const int a = idx / ow, b = idx % ow;
                if (b > a) continue;
                F acc = 0;
                for (int i = 0; i < NOBS(); ++i) {
                    const F* J = ws + oLJO() + (k * NOBS() + i) * ow;
                    acc += wrow[i] * J[a] * J[b];
                }
                sM[(nq + a) * ld + nq + b] += acc;
            }
            tsync();
        }
    }

    // ------------------------------------------------------------ force block
    // D^-1 of stage k from the side records (force box rows, friction pyramid rows) -> sD; q = D^-1 m_f -> sFq
    // (m_f = force part of the stage gradient in sVec).
    __device__ void force_block_D(int k) {
        const int nq = NQ(), nf = NF();
        const R base = PR.dt * PR.fw + PR.reg_input;
        if (nf == 1) {
            const R eps = row_eps(0);
            for (int c = lane; c < NC(); c += kTS) sD[c] = rinv(base + box_weight_r(recs_tl(k)[nq + c], eps));
        } else {
            const R eps0 = row_eps(0), eps2 = row_eps(2);
            const int nbx = NBOXU() + NX();
            for (int c = lane; c < NC(); c += kTS) {
                // symmetric 3 x 3: m00 m10 m11 m20 m21 m22
                R m00 = base + box_weight_r(recs_tl(k)[nq + 3 * c], eps0), m11 = base + box_weight_r(recs_tl(k)[nq + 3 * c + 1], eps0),
                  m22 = base + box_weight_r(recs_tl(k)[nq + 3 * c + 2], eps0), m10 = 0, m20 = 0, m21 = 0;
                for (int which = 0; which < 5; ++which) {
                    const R w = one_weight_r(recs_tl(k)[nbx + 5 * c + which], eps2);
                    const V3<R> a = fric_coeff(c, which);
                    m00 += w * a.x * a.x;
                    m10 += w * a.y * a.x;
                    m11 += w * a.y * a.y;
                    m20 += w * a.z * a.x;
                    m21 += w * a.z * a.y;
                    m22 += w * a.z * a.z;
                }
                // inverse by cofactors (SPD)
                const R c00 = m11 * m22 - m21 * m21, c10 = m20 * m21 - m10 * m22, c20 = m10 * m21 - m20 * m11;
                const R det = m00 * c00 + m10 * c10 + m20 * c20;
                const R id = rinv(det);
                R* o = sD + 9 * c;
                o[0] = c00 * id;
                o[1] = o[3] = c10 * id;
                o[2] = o[6] = c20 * id;
                o[4] = (m00 * m22 - m20 * m20) * id;
                o[5] = o[7] = (m10 * m20 - m00 * m21) * id;
                o[8] = (m00 * m11 - m10 * m10) * id;
            }
        }
        tsync();
    }
    // q = D^-1 g_f for the force part of a stage gradient; Dm = D^-1 of the stage (shared or global)
    __device__ __forceinline__ void force_q(const R* Dm, const R* gf, R* q) const {
        const int nf = NF();
        if (nf == 1) {
            for (int c = lane; c < NC(); c += kTS) q[c] = Dm[c] * gf[c];
        } else {
            for (int c = lane; c < NC(); c += kTS) {
                const R* o = Dm + 9 * c;
                const R g0 = gf[3 * c], g1 = gf[3 * c + 1], g2 = gf[3 * c + 2];
                q[3 * c] = o[0] * g0 + o[1] * g1 + o[2] * g2;
                q[3 * c + 1] = o[3] * g0 + o[4] * g1 + o[5] * g2;
                q[3 * c + 2] = o[6] * g0 + o[7] * g1 + o[8] * g2;
            }
        }
    }
    // S = R^-1 + Df D^-1 Df' per group -> sS (lower triangles, dense ng x ng per group), then its Cholesky factor
    // and the inverse of that factor (lower) in place.  Returns false on a non-positive / non-finite pivot.
    // Groups of one body (ng = 6: every BASELINE configuration but the stacked one): lane g builds, factors and
    // inverts the 6 x 6 block of body g entirely in registers (fully unrolled, no barriers).
    __device__ bool force_block_S6(int k) {
        const int ngrp = NGRP(), nf = NF();
        const F* Dc = sDFC;
        bool ok = true;
        for (int g = lane; g < ngrp; g += kTS) {
            R s[21];   // lower triangle, row-major: (a, b) at a (a + 1) / 2 + b
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) s[a * (a + 1) / 2 + b] = R(0);
#pragma unroll
            for (int a = 0; a < 6; ++a) s[a * (a + 1) / 2 + a] = rho_inv_stage(k, 6 * g + a);
            for (int l = P.bc_start[g]; l < P.bc_start[g + 1]; ++l) {
                const int cs = P.bc_list[l], c = cs >> 1;
                const F* n = Dc + cs * 6 * nf;
                if (nf == 1) {
                    const R d = sD[c];
                    R na[6];
#pragma unroll
                    for (int a = 0; a < 6; ++a) na[a] = R(n[a]);
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const R da = d * na[a];
#pragma unroll
                        for (int b = 0; b <= a; ++b) s[a * (a + 1) / 2 + b] += da * na[b];
                    }
                } else {
                    const R* o = sD + 9 * c;
                    R na[6][3], ua[6][3];
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) na[a][i] = R(n[a * 3 + i]);
#pragma unroll
                        for (int i = 0; i < 3; ++i) ua[a][i] = o[3 * i] * na[a][0] + o[3 * i + 1] * na[a][1] + o[3 * i + 2] * na[a][2];
                    }
#pragma unroll
                    for (int a = 0; a < 6; ++a)
#pragma unroll
                        for (int b = 0; b <= a; ++b)
                            s[a * (a + 1) / 2 + b] += ua[a][0] * na[b][0] + ua[a][1] * na[b][1] + ua[a][2] * na[b][2];
                }
            }
            // Cholesky; id[j] = 1 / L_jj
            R id[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const R d = s[j * (j + 1) / 2 + j];
                if (!(d > R(0)) || !(d < R(1e300))) ok = false;
                id[j] = rsqrt(d);
#pragma unroll
                for (int i = j + 1; i < 6; ++i) s[i * (i + 1) / 2 + j] *= id[j];
#pragma unroll
                for (int i = j + 1; i < 6; ++i)
#pragma unroll
                    for (int l = j + 1; l <= i; ++l) s[i * (i + 1) / 2 + l] -= s[i * (i + 1) / 2 + j] * s[l * (l + 1) / 2 + j];
            }
            // inverse of the factor in place, column by column (column c of L^-1 needs the original columns > c only)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                s[c * (c + 1) / 2 + c] = id[c];
#pragma unroll
                for (int i = c + 1; i < 6; ++i) {
                    R acc = s[i * (i + 1) / 2 + c] * id[c];
#pragma unroll
                    for (int l = c + 1; l < i; ++l) acc += s[i * (i + 1) / 2 + l] * s[l * (l + 1) / 2 + c];
                    s[i * (i + 1) / 2 + c] = -acc * id[i];
                }
            }
            R* Sg = sS + g * 36;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b) Sg[a * 6 + b] = b <= a ? s[a * (a + 1) / 2 + b] : R(0);
        }
        tsync();
        return __all_sync(FULL, ok);
    }
    // Bodies that share contacts, compile-time size (cfg3: 18 rows): Cholesky with lane = row, the row in registers
    // and the pivot column travelling by warp shuffle (no barriers, no index arithmetic).  sS then holds L itself
    // (lower, diagonal INVERTED), not its inverse: the products with L^-1 / L^-T become substitutions.
    static constexpr bool kPanelS = D::kStatic && D::ng > 6 && D::ng <= 32;
    __device__ bool force_block_S(int k) {
        const int ng = NG(), ngrp = NGRP(), nf = NF();
        if (ng == 6) return force_block_S6(k);
        const F* Dc = sDFC;
        // U = D^-1 N' per contact and side (scratch): S_ab = sum_c N[a] . U[b] then costs nf multiply-adds per contact
        for (int e = lane; e < NC() * 2; e += kTS) {
            const int c = e >> 1;
            if (((e & 1) == 0 ? P.cb2[c] : P.cb1[c]) < 0) continue;
            const F* n = Dc + (e * 6) * nf;
            R* u = sUS + (e * 6) * nf;
            if (nf == 1) {
                for (int rr = 0; rr < 6; ++rr) u[rr] = sD[c] * R(n[rr]);
            } else {
                const R* o = sD + 9 * c;
                for (int rr = 0; rr < 6; ++rr) {
                    const R n0 = R(n[3 * rr]), n1 = R(n[3 * rr + 1]), n2 = R(n[3 * rr + 2]);
                    u[3 * rr] = o[0] * n0 + o[1] * n1 + o[2] * n2;
                    u[3 * rr + 1] = o[3] * n0 + o[4] * n1 + o[5] * n2;
                    u[3 * rr + 2] = o[6] * n0 + o[7] * n1 + o[8] * n2;
                }
            }
        }
        tsync();
        const int tri = ng * (ng + 1) / 2;
        for (int e = lane; e < ngrp * tri; e += kTS) {
            const int g = e / tri, t = e - g * tri;
            // (a, b) of the t-th lower-triangle entry, row by row
            int a = int((sqrtf(float(8 * t + 1)) - 1.0f) * 0.5f);
            while ((a + 1) * (a + 2) / 2 <= t) ++a;
            while (a * (a + 1) / 2 > t) --a;
            const int b = t - a * (a + 1) / 2;
            const int ra = g * ng + a, rb = g * ng + b;
            const int ba = ra / 6, bb = rb / 6, ia = ra - 6 * ba, ib = rb - 6 * bb;
            R acc = R(0);
            if (a == b) acc = rho_inv_stage(k, ra);
#pragma unroll 1
            for (int l = P.bc_start[ba]; l < P.bc_start[ba + 1]; ++l) {
                const int cs = P.bc_list[l], c = cs >> 1;
                // the side of contact c that touches body bb (if any)
                int cs_b = -1;
                if (bb == ba) cs_b = cs;
                else if (P.cb2[c] == bb) cs_b = 2 * c;
                else if (P.cb1[c] == bb) cs_b = 2 * c + 1;
                if (cs_b < 0) continue;
                const F* na = Dc + (cs * 6 + ia) * nf;
                const R* ub = sUS + (cs_b * 6 + ib) * nf;
                if (nf == 1) acc += R(na[0]) * ub[0];
                else acc += R(na[0]) * ub[0] + R(na[1]) * ub[1] + R(na[2]) * ub[2];
            }
            sS[g * ng * ng + a * ng + b] = acc;
            if constexpr (kPanelS) sS[g * ng * ng + b * ng + a] = acc;   // the panel reads whole rows
        }
        tsync();
        bool ok = true;
        if constexpr (kPanelS) {
            constexpr int NG_ = D::ng;
            R row[NG_];
            const int a = lane < NG_ ? lane : NG_ - 1;    // lanes beyond the last row shadow it (their results are dropped)
#pragma unroll
            for (int b = 0; b < NG_; ++b) row[b] = sS[a * NG_ + b];
#pragma unroll
            for (int j = 0; j < NG_; ++j) {
                const R d = __shfl_sync(FULL, row[j], j);
                if (!(d > R(0)) || !(d < R(1e300))) ok = false;
                const R inv = rsqrt(d);
                const R lij = (lane > j) ? row[j] * inv : R(0);
#pragma unroll
                for (int c = j + 1; c < NG_; ++c) {
                    const R lcj = __shfl_sync(FULL, lij, c);
                    row[c] -= lij * lcj;
                }
                row[j] = (lane > j) ? lij : (lane == j ? inv : R(0));
            }
            tsync();
            if (lane < NG_) {
#pragma unroll
                for (int b = 0; b < NG_; ++b) sS[lane * NG_ + b] = row[b];
            }
            tsync();
            return __all_sync(FULL, ok);
        } else {
        // run-time size: Cholesky in shared memory, all groups at once: lanes over (group, row) pairs for the column
        // scaling and over (group, i, l) for the trailing update
        for (int j = 0; j < ng; ++j) {
            const int nrow_j = ng - 1 - j;
            // pivots: every lane reads the pivot of the group it works on
            for (int e = lane; e < ngrp * nrow_j; e += kTS) {
                const int g = e / nrow_j, i = j + 1 + e % nrow_j;
                R* Sg = sS + g * ng * ng;
                const R d = Sg[j * ng + j];
                if (!(d > R(0)) || !(d < R(1e300))) ok = false;
                Sg[i * ng + j] *= rsqrt(d);
            }
            tsync();
            const int ntr = nrow_j * (nrow_j + 1) / 2;
            for (int e = lane; e < ngrp * ntr; e += kTS) {
                const int g = e / ntr, t = e - g * ntr;
                int a = int((sqrtf(float(8 * t + 1)) - 1.0f) * 0.5f);
                while ((a + 1) * (a + 2) / 2 <= t) ++a;
                while (a * (a + 1) / 2 > t) --a;
                const int b = t - a * (a + 1) / 2;
                R* Sg = sS + g * ng * ng;
                const int i = j + 1 + a, l = j + 1 + b;
                Sg[i * ng + l] -= Sg[i * ng + j] * Sg[l * ng + j];
            }
            tsync();
            for (int g = lane; g < ngrp; g += kTS) {
                R* Sg = sS + g * ng * ng;
                const R d = Sg[j * ng + j];
                if (!(d > R(0)) || !(d < R(1e300))) ok = false;
                Sg[j * ng + j] = sqrt(d);
            }
            tsync();
        }
        // inverse of the factor: lane (g, c) computes column c of L^-1 by forward substitution into the UPPER
        // triangle of the same block (row c, columns c..ng-1 hold column c of L^-1), then the block is rewritten as
        // the lower-triangular L^-1
        for (int e = lane; e < ngrp * ng; e += kTS) {
            const int g = e / ng, c = e - g * ng;
            R* Sg = sS + g * ng * ng;
            Sg[c * ng + c] = R(1) / Sg[c * ng + c];   // the diagonal now holds 1 / L_cc = (L^-1)_cc
        }
        tsync();
        for (int e = lane; e < ngrp * ng; e += kTS) {
            const int g = e / ng, c = e - g * ng;
            R* Sg = sS + g * ng * ng;
            for (int i = c + 1; i < ng; ++i) {
                R s = Sg[i * ng + c] * Sg[c * ng + c];              // L_ic x_c
                for (int l = c + 1; l < i; ++l) s += Sg[i * ng + l] * Sg[c * ng + l];   // L_il x_l (x_l kept at [c][l])
                Sg[c * ng + i] = -s * Sg[i * ng + i];              // diagonal already inverted
            }
        }
        tsync();
        // move the columns of L^-1 from the upper triangle to the lower one: (L^-1)_ic = upper[c][i]
        for (int e = lane; e < ngrp * ng * ng; e += kTS) {
            const int g = e / (ng * ng), t = e - g * ng * ng, i = t / ng, c = t % ng;
            if (i > c) sS[g * ng * ng + i * ng + c] = sS[g * ng * ng + c * ng + i];
        }
        tsync();
        return __all_sync(FULL, ok);
        }
    }
    // y = L^-1 x per group (rows over lanes); Lm = L^-1 blocks (lower), shared or global
    __device__ __forceinline__ void force_Linv_mul(const R* Lm, const R* x, R* y) const {
        if constexpr (kPanelS) {   // Lm = L (diagonal inverted): forward substitution, lane = row, x_j broadcast by shuffle
            constexpr int NG_ = D::ng;
            R v = lane < NG_ ? x[lane] : R(0);
            const R* row = Lm + (lane < NG_ ? lane : 0) * NG_;
#pragma unroll
            for (int j = 0; j < NG_; ++j) {
                const R xj = __shfl_sync(FULL, v, j) * Lm[j * NG_ + j];
                if (lane == j) v = xj;
                else if (lane > j && lane < NG_) v -= row[j] * xj;
            }
            if (lane < NG_) y[lane] = v;
            return;
        }
        const int ng = NG();
        for (int r = lane; r < NEQ(); r += kTS) {
            const int g = r / ng, a = r - g * ng;
            const R* row = Lm + g * ng * ng + a * ng;
            R s = 0;
            for (int l = 0; l <= a; ++l) s += row[l] * x[g * ng + l];
            y[r] = s;
        }
    }
    // y = L^-T x per group
    __device__ __forceinline__ void force_LinvT_mul(const R* Lm, const R* x, R* y) const {
        if constexpr (kPanelS) {   // back substitution with L' (row a of L read by the lanes below a)
            constexpr int NG_ = D::ng;
            R v = lane < NG_ ? x[lane] : R(0);
#pragma unroll
            for (int a = NG_ - 1; a >= 0; --a) {
                const R ya = __shfl_sync(FULL, v, a) * Lm[a * NG_ + a];
                if (lane == a) v = ya;
                else if (lane < a) v -= Lm[a * NG_ + lane] * ya;
            }
            if (lane < NG_) y[lane] = v;
            return;
        }
        const int ng = NG();
        for (int r = lane; r < NEQ(); r += kTS) {
            const int g = r / ng, a = r - g * ng;
            const R* blk = Lm + g * ng * ng;
            R s = 0;
            for (int l = a; l < ng; ++l) s += blk[l * ng + a] * x[g * ng + l];
            y[r] = s;
        }
    }
    // Force block of stage k < N in the factor pass.  On entry: sC = C rows, sVec = stage gradient [j; f; x] (double),
    // sFv = v = e + y / rho, sD = D^-1.  On exit the shared-memory bundle holds G (in place of C), L^-1, D^-1,
    // g_lambda and q, sGf = g_lambda in F, and the bundle is stored for the later passes.
    __device__ bool force_block_factor(int k) {
        const int nq = NQ(), nx = NX(), ne = NEQ(), ng = NG();
        force_q(sD, sVec + nq, sFq);
        const bool ok = force_block_S(k);
        tsync();
        // rhs = v - Df q  (in place in sFv)
        for (int r = lane; r < ne; r += kTS) sFv[r] -= eq_force_dot(r, sFq);
        tsync();
        force_Linv_mul(sS, sFv, sFg);
        if constexpr (kPanelS) {
            // G = L^-1 C by forward substitution, lanes over columns, the column in registers
            constexpr int NG_ = D::ng;
            if (lane < nx) {
                R gcol[NG_];
#pragma unroll
                for (int a = 0; a < NG_; ++a) {
                    R sacc = R(cC[a * nx + lane]);
#pragma unroll
                    for (int l = 0; l < a; ++l) sacc -= sS[a * NG_ + l] * gcol[l];
                    gcol[a] = sacc * sS[a * NG_ + a];
                }
#pragma unroll
                for (int a = 0; a < NG_; ++a) sC[a * nx + lane] = F(gcol[a]);
            }
        } else
        // G = L^-1 C, bottom row first (in place when the C rows sit in sC: row a needs the original rows <= a only);
        // lanes over columns
        for (int j = lane; j < nx; j += kTS) {
            for (int g = 0; g < NGRP(); ++g) {
                const R* blk = sS + g * ng * ng;
                const F* src = cC + (g * ng) * nx + j;
                F* col = sC + (g * ng) * nx + j;
                for (int a = ng - 1; a >= 0; --a) {
                    R s = 0;
                    for (int l = 0; l <= a; ++l) s += blk[a * ng + l] * R(src[l * nx]);
                    col[a * nx] = F(s);
                }
            }
        }
        tsync();
        for (int r = lane; r < ne; r += kTS) sGf[r] = F(sFg[r]);
        // keep for the corrector / forward passes: one coalesced copy of the bundle
        {
            const int4* src = reinterpret_cast<const int4*>(sFB);
            int4* dst = reinterpret_cast<int4*>(bundle(k));
            const int n16 = obsize() * int(sizeof(F)) / 16;
            for (int i = lane; i < n16; i += kTS) dst[i] = src[i];
        }
        tsync();
        return ok;
    }
    // Force step of stage k for the state step dx (sDst + nu): lambda = L^-T (g_lambda + G dx),
    // df = -q - D^-1 Df' lambda  -> sDst[nq, nu).  The bundle of the stage is in shared memory.
    __device__ void force_block_step() {
        const int nq = NQ(), nx = NX(), nu = NU(), ne = NEQ(), nf = NF();
        const F* dx = sDst + nu;
        if (ne <= 8) {
            const int r = lane >> 2, part = lane & 3;
            R s = R(0);
            if (r < ne) {
                const F* g = sC + r * nx;
                for (int j = part; j < nx; j += 4) s += R(g[j]) * R(dx[j]);
                if (part == 0) s += sFg[r];
            }
            s += __shfl_xor_sync(FULL, s, 1);
            s += __shfl_xor_sync(FULL, s, 2);
            if (r < ne && part == 0) sFv[r] = s;
        } else {
            for (int r = lane; r < ne; r += kTS) {
                const F* g = sC + r * nx;
                R s = sFg[r];
                for (int j = 0; j < nx; ++j) s += R(g[j]) * R(dx[j]);
                sFv[r] = s;
            }
        }
        tsync();
        force_LinvT_mul(sS, sFv, sFl);
        tsync();
        const F* Dc = sDFC;
        for (int c = lane; c < NC(); c += kTS) {
            R u[3] = {0, 0, 0};
            for (int side = 0; side < 2; ++side) {
                const int b = side == 0 ? P.cb2[c] : P.cb1[c];
                if (b < 0) continue;
                const F* blk = Dc + ((2 * c + side) * 6) * nf;
                for (int rr = 0; rr < 6; ++rr) {
                    const R lm = sFl[6 * b + rr];
                    for (int i = 0; i < nf; ++i) u[i] += R(blk[rr * nf + i]) * lm;
                }
            }
            if (nf == 1) sDst[nq + c] = F(-sFq[c] - sD[c] * u[0]);
            else {
                const R* o = sD + 9 * c;
                sDst[nq + 3 * c] = F(-sFq[3 * c] - (o[0] * u[0] + o[1] * u[1] + o[2] * u[2]));
                sDst[nq + 3 * c + 1] = F(-sFq[3 * c + 1] - (o[3] * u[0] + o[4] * u[1] + o[5] * u[2]));
                sDst[nq + 3 * c + 2] = F(-sFq[3 * c + 2] - (o[6] * u[0] + o[7] * u[1] + o[8] * u[2]));
            }
        }
        tsync();
    }
    // the force bundle of stage k from the workspace into shared memory: synchronously, or issued as cp.async
    // (no commit) for the staged kernels
    __device__ void load_force_block(int k) {
        if (NFC() == 0 || k < 0 || k >= NN()) return;
        const int4* __restrict__ src = reinterpret_cast<const int4*>(bundle(k));
        int4* dst = reinterpret_cast<int4*>(sFB);
        const int n16 = obsize() * int(sizeof(F)) / 16;
        for (int i = lane; i < n16; i += kTS) dst[i] = src[i];
        tsync();
    }
    __device__ __forceinline__ void fb_issue(int k) const {
        if constexpr (kStageFB) {
            if (NFC() == 0 || k < 0 || k >= NN()) return;
            cp_async_bytes(sFB, bundle(k), obsize() * int(sizeof(F)));
        }
    }

    // Blocked factorisation of the AUGMENTED reduced stage matrix [M m; m' .]:
    //   panel   [L; Y; w'] = [M; m'][:, 0:nq] L^{-T}   right-looking, lane = matrix row (rows in registers, the pivot
    //                                                  column travels by warp shuffle), diagonal stored INVERTED;
    //                                                  the gradient rides along as row nr, so the forward
    //                                                  substitution w = L^{-1} m_j costs nothing extra;
    //   Schur   [P p] = [Mxx m_x] + [G g]'[G g] - Y [Y; w']'   8 x 4 lane grid, TR x TC accumulator tile per lane; the
    //                                                  spare tile column carries p = m_x + G'g_lambda - Y w.  The rows
    //                                                  of [G | g_lambda] (force block, shared memory sC) enter here, so
    //                                                  C' S^-1 C is never written to the stage matrix.  P goes
    //                                                  straight to sP as the full symmetric cost-to-go Hessian, p to sPv.
    // The factor block is stored to the workspace column-major from the panel registers (coalesced).
    // On entry vec = reduced stage gradient [m_j; m_x]; on exit vec[0, nq) = w.
    __device__ bool stage_factor_blocked(F* vec, F* Fg, int n_g) {
        constexpr int NU_ = D::nq, NX_ = D::nx, NZ_ = NU_ + NX_, AUG = NZ_;
        constexpr int RR = (NZ_ + 1 + kTS - 1) / kTS;
        constexpr int TR = (NX_ + 7) / 8, TC = (NX_ + 1 + 3) / 4;
        static_assert(NU_ <= 16 && RR <= 2 && 4 * TC > NX_, "blocked factorisation is for small stage matrices");
        constexpr int ld = LDM();
        F row[RR][NU_];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const int i = lane + kTS * r;
            const F* src = (i == AUG) ? vec : sM + min(i, NZ_ - 1) * ld;
#pragma unroll
            for (int c = 0; c < NU_; ++c) row[r][c] = src[c];
        }
        bool ok = true;
#pragma unroll
        for (int j = 0; j < NU_; ++j) {
            // The exact pivot is bounded below by reg_input (M_jj >= reg_input I); one that roundoff pushed under
            // the bound is raised to it, which keeps the fp32 factorisation finite (the Newton step is then inexact
            // and the interior-point iteration corrects it).  A NaN pivot still fails the solve.
            F d = __shfl_sync(FULL, row[0][j], j);
            if (d != d) ok = false;
            if (!(d > C.reg_input)) d = C.reg_input;
            const F inv = rsqrt(d);
            F lij[RR];
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const int i = lane + kTS * r;
                lij[r] = (i > j) ? row[r][j] * inv : F(0);
            }
#pragma unroll
            for (int c = j + 1; c < NU_; ++c) {
                const F lcj = __shfl_sync(FULL, lij[0], c);
#pragma unroll
                for (int r = 0; r < RR; ++r) row[r][c] -= lij[r] * lcj;
            }
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const int i = lane + kTS * r;
                row[r][j] = (i > j) ? lij[r] : (i == j ? inv : F(0));
            }
        }
        tsync();   // lanes beyond the last row read a clamped (real) row above
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const int i = lane + kTS * r;
            if (i < NZ_) {
#pragma unroll
                for (int c = 0; c < NU_; ++c) {
                    if (i >= NU_) sM[i * ld + c] = row[r][c];   // Y feeds the Schur update
                    Fg[c * (NZ_ + 1) + i] = row[r][c];          // factor block -> workspace (column-major)
                }
            } else if (i == AUG) {
#pragma unroll
                for (int c = 0; c < NU_; ++c) vec[c] = row[r][c];  // w = L^{-1} m_j
            }
        }
        tsync();
        // Schur complement tile of lane (a, b): rows a*TR.., columns b*TC.. of the state block; column NX_ = p
        const int a = lane >> 2, b = lane & 3;
        int ri[TR];
        const F* ycp[TC];
        const F* gcp[TC];   // column c of [G | g_lambda]: G (row stride nx) or the F copy of g_lambda (stride 1)
        int gst[TC];
#pragma unroll
        for (int ii = 0; ii < TR; ++ii) ri[ii] = min(a * TR + ii, NX_ - 1);
        F acc[TR][TC];
#pragma unroll
        for (int cc = 0; cc < TC; ++cc) {
            const int c = b * TC + cc;
            const int cl = min(c, NX_ - 1);
            ycp[cc] = (c == NX_) ? vec : sM + (NU_ + cl) * ld;
            gcp[cc] = (c >= NX_) ? sGf : sC + cl;
            gst[cc] = (c >= NX_) ? 1 : NX_;
#pragma unroll
            for (int ii = 0; ii < TR; ++ii) {
                const int hi = max(ri[ii], cl), lo = min(ri[ii], cl);
                acc[ii][cc] = (c == NX_) ? vec[NU_ + ri[ii]] : sM[(NU_ + hi) * ld + NU_ + lo];
            }
        }
#pragma unroll 1
        for (int m = 0; m < NU_; ++m) {   // rolled: 39 instructions that stay in the instruction cache
            F yr[TR], yc[TC];
#pragma unroll
            for (int ii = 0; ii < TR; ++ii) yr[ii] = sM[(NU_ + ri[ii]) * ld + m];
#pragma unroll
            for (int cc = 0; cc < TC; ++cc) yc[cc] = ycp[cc][m];
#pragma unroll
            for (int ii = 0; ii < TR; ++ii)
#pragma unroll
                for (int cc = 0; cc < TC; ++cc) acc[ii][cc] -= yr[ii] * yc[cc];
        }
#pragma unroll 1
        for (int m = 0; m < n_g; ++m) {   // + [G g]'[G g]: the force block's contribution
            const F* gr = sC + m * NX_;
            F yr[TR], yc[TC];
#pragma unroll
            for (int ii = 0; ii < TR; ++ii) yr[ii] = gr[ri[ii]];
#pragma unroll
            for (int cc = 0; cc < TC; ++cc) yc[cc] = gcp[cc][m * gst[cc]];
#pragma unroll
            for (int ii = 0; ii < TR; ++ii)
#pragma unroll
                for (int cc = 0; cc < TC; ++cc) acc[ii][cc] += yr[ii] * yc[cc];
        }
#pragma unroll
        for (int ii = 0; ii < TR; ++ii)
#pragma unroll
            for (int cc = 0; cc < TC; ++cc) {
                const int i = a * TR + ii, c = b * TC + cc;
                if (i < NX_ && c < NX_) sP[i * NX_ + c] = acc[ii][cc];
                else if (i < NX_ && c == NX_) sPv[i] = acc[ii][cc];
            }
        tsync();
        return ok;
    }
    // cost-to-go Hessian = state block of sM, expanded to the full symmetric matrix (terminal stage)
    __device__ __forceinline__ void copy_cost_to_go() {
        constexpr int nq = D::nq, nx = D::nx, ld = LDM();
        for (int idx = lane; idx < nx * nx; idx += kTS) {
            const int i = idx / nx, j = idx % nx;
            sP[idx] = (j <= i) ? sM[(nq + i) * ld + nq + j] : sM[(nq + j) * ld + nq + i];
        }
        tsync();
    }

    // Stage gradient of the barrier Lagrangian at the current iterate (double, stage layout [j; f; x] in sVec):
    //   H z + g  +  sum_sides sgn a [ -lam + (rc + lam rd)/(t + eps lam) ]   (+ terminal rows a (rho e + y) at k = N)
    // with rc = t lam - target (+ dt_aff dlam_aff in the corrector).  The object-dynamics rows of k < N are NOT in
    // it: their values e + y / rho go to sFv and meet the forces in the force block.
    // Every row family accumulates into distinct entries per lane (no atomics).
    __device__ void stage_gradient(int k, bool corrector, F mu_target) {
        const int nq = NQ(), nu = NU(), nx = NX(), nz = NZ();
        const R dt = PR.dt;
        const R* zk = st_z(k);
        R* vec = sVec;
        // cost part (zero at the terminal stage)
        if (k < NN()) {
            const F* x = st_x(k);
            const F* u = st_u(k);
            const F* Jp = st_jp(k);
            // e = Jp dq + r - r_d, reduced over the warp
            R e3[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const R part = lane < nq ? R(Jp[c * nq + lane]) * zk[nu + lane] : R(0);
                e3[c] = tsum(part) + wsr<R>(oLR())[3 * k + c] - R(target[3 * k + c]);
            }
            // orientation error at the QP iterate: e_o + Jq dq
            R eo3[3] = {R(0), R(0), R(0)};
            const F* Jq = ws + oLJQ() + k * 3 * nq;
            if (ORI()) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const R part = lane < nq ? R(Jq[c * nq + lane]) * zk[nu + lane] : R(0);
                    eo3[c] = tsum(part) + wsr<R>(oLRO())[3 * k + c];
                }
            }
            // inertial-alignment residual at the QP iterate: e + Je dx
            R ea[2] = {R(0), R(0)};
            const F* Ja = ws + oLJA() + 2 * k * nx;
            if (IALIGN()) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    R part = R(0);
                    for (int j = lane; j < nx; j += kTS) part += R(Ja[r * nx + j]) * zk[nu + j];
                    ea[r] = tsum(part) + R(ws[oLIA() + 2 * k + r]);
                }
            }
            for (int i = lane; i < nz; i += kTS) {
                R g;
                if (i < nq) g = dt * PR.Rd[i] * (R(u[i]) + zk[i]) + PR.reg_input * zk[i];
                else if (i < nu) g = dt * PR.fw * (R(u[i]) + zk[i]) + PR.reg_input * zk[i];
                else {
                    const int xi = i - nu;
                    g = dt * PR.Qd[xi] * (R(x[xi]) + zk[i] - PR.xd[xi]);
                    if (xi < nq)
                        g += dt * (PR.Wd[0] * R(Jp[xi]) * e3[0] + PR.Wd[1] * R(Jp[nq + xi]) * e3[1] + PR.Wd[2] * R(Jp[2 * nq + xi]) * e3[2]);
                    if (ORI() && xi < nq)
                        g += dt * (PR.Wo[0] * R(Jq[xi]) * eo3[0] + PR.Wo[1] * R(Jq[nq + xi]) * eo3[1] + PR.Wo[2] * R(Jq[2 * nq + xi]) * eo3[2]);
                    if (IALIGN()) g += dt * PR.ia_w * (R(Ja[xi]) * ea[0] + R(Ja[nx + xi]) * ea[1]);
                }
                vec[i] = g;
            }
        } else {
            for (int i = lane; i < nz; i += kTS) vec[i] = R(0);
        }
        tsync();
        // box rows: one entry each
        for (int r = lane; r < NBOXU() + nx; r += kTS) {
            const int fam = r < NBOXU() ? 0 : 1;
            if (!row_valid(k, fam)) continue;
            const int m = fam == 0 ? r : nu + (r - NBOXU());
            R lb, ub;
            const R val = zk[m];
            if (fam == 0) {
                const R uu = R(st_u(k)[r]);
                lb = (r < nq ? PR.ulb[r] : PR.flb) - uu;
                ub = (r < nq ? PR.uub[r] : PR.fub) - uu;
            } else {
                const int i = r - NBOXU();
                const R xx = R(st_x(k)[i]);
                lb = PR.xlb[i] - xx;
                ub = PR.xub[i] - xx;
            }
            const R eps = row_eps(fam);
            const QuadR q = recs_tl(k)[r];
            F c0 = F(0), c1 = F(0);
            if (corrector) {
                const QuadF dd = recs_dd(k)[r];
                c0 = dd.v[0] * dd.v[2];
                c1 = dd.v[1] * dd.v[3];
            }
            vec[m] += side_coef(q.v[0], q.v[2], val - lb, eps, mu_target, c0) -
                      side_coef(q.v[1], q.v[3], ub - val, eps, mu_target, c1);
        }
        tsync();
        // equality rows
        if (k < NN()) {
            eq_values_stage(k, zk, sFv);
            if (!C.soft_poly) {   // hard rows carry multipliers: v = e + y / rho
                tsync();
                const R* y = y_eq(k);
                for (int i = lane; i < NEQ(); i += kTS) sFv[i] += y[i] * rho_inv_stage(k, i);
            }
        } else {
            // terminal rows: m_i = rho e_i + y_i; unit rows touch distinct entries, the three dense rows go through
            // the scratch
            const R* rho = rho_eq(k);
            const R* y = y_eq(k);
            R* mrow = sScr;
            for (int i = lane; i < NTERM(); i += kTS) {
                const R m = rho[i] * eq_value_term(i, zk) + y[i];
                if (i < 3) mrow[i] = m;
                else vec[nu + nq + (i - 3)] += m;
            }
            tsync();
            if (lane < nq) {
                const F* Jp = ws + oLJP() + k * 3 * nq;
                vec[nu + lane] -= mrow[0] * R(Jp[lane]) + mrow[1] * R(Jp[nq + lane]) + mrow[2] * R(Jp[2 * nq + lane]);
            }
        }
        tsync();
        const int nbx = NBOXU() + nx;
        if (NFRIC() > 0 && k < NN()) {
            // one lane per contact: five pyramid rows -> three force entries
            const R eps = row_eps(2);
            for (int c = lane; c < NC(); c += kTS) {
                const F* f = st_u(k) + nq + 3 * c;
                const R* df = zk + nq + 3 * c;
                const R f0 = R(f[0]) + df[0], f1 = R(f[1]) + df[1], f2 = R(f[2]) + df[2];
                R g0 = 0, g1 = 0, g2 = 0;
                for (int which = 0; which < 5; ++which) {
                    const int r = nbx + 5 * c + which;
                    const V3<R> a = fric_coeff(c, which);
                    const R val = a.x * f0 + a.y * f1 + a.z * f2;
                    const QuadR q = recs_tl(k)[r];
                    F corr = F(0);
                    if (corrector) {
                        const QuadF dd = recs_dd(k)[r];
                        corr = dd.v[0] * dd.v[2];
                    }
                    const R cf = side_coef(q.v[0], q.v[2], val, eps, mu_target, corr);
                    g0 += cf * a.x;
                    g1 += cf * a.y;
                    g2 += cf * a.z;
                }
                vec[nq + 3 * c] += g0;
                vec[nq + 3 * c + 1] += g1;
                vec[nq + 3 * c + 2] += g2;
            }
            tsync();
        }
        if (NOBS() > 0 && k >= 1 && k < NN()) {
            const R eps = row_eps(3);
            R* crow = sScr;
            for (int i = lane; i < NOBS(); i += kTS) {
                const int r = nbx + NFRIC() + i;
                const F* J = ws + oLJO() + (k * NOBS() + i) * OBSW();
                R val = R(ws[oLHO() + k * NOBS() + i]);
                for (int j = 0; j < OBSW(); ++j) val += R(J[j]) * zk[nu + j];
                const QuadR q = recs_tl(k)[r];
                F corr = F(0);
                if (corrector) {
                    const QuadF dd = recs_dd(k)[r];
                    corr = dd.v[0] * dd.v[2];
                }
                crow[i] = side_coef(q.v[0], q.v[2], val, eps, mu_target, corr);
            }
            tsync();
            if (lane < OBSW()) {
                R acc = 0;
                for (int i = 0; i < NOBS(); ++i) acc += crow[i] * R(ws[oLJO() + (k * NOBS() + i) * OBSW() + lane]);
                vec[nu + lane] += acc;
            }
            tsync();
        }
    }

    // Factor-block staging: the Riccati block of the NEXT stage is fetched with cp.async (LDGSTS, generic proxy — no
    // proxy fence against the plain stores of the factor pass) into the other half of the idle stage-matrix
    // buffer while the current stage is processed.
    __device__ __forceinline__ void fac_issue(int k, int buf) {
        if (k < 0 || k >= NN()) return;
        cp_async_bytes(sM + buf * FSTRIDE(), ws + oFAC() + k * FSTRIDE(), FSTRIDE() * int(sizeof(F)));
    }
    // reduced right-hand side [m_j; m_x + G' g_lambda] from the stage gradient (sVec) — G read from `Gsrc`
    // (row stride gs), g_lambda from gl; k = N: no force block
    __device__ __forceinline__ void reduced_rhs(int k, const F* Gsrc, int gs, const R* gl) {
        const int nq = NQ(), nx = NX(), nu = NU();
        for (int i = lane; i < nq; i += kTS) sRv[i] = F(sVec[i]);
        for (int j = lane; j < nx; j += kTS) {
            R s = sVec[nu + j];
            if (k < NN() && Gsrc != nullptr)
                for (int r = 0; r < NEQ(); ++r) s += R(Gsrc[r * gs + j]) * gl[r];
            sRv[nq + j] = F(s);
        }
        tsync();
    }

    // ---------------------------------------------------------------- fused IPM passes
    // Pass A (backward): per stage form the predictor gradient (sigma = 0), keep it in GP for the corrector,
    // build the reduced stage matrix, eliminate the forces, factor and do the backward vector step with the
    // factor still in shared memory.
    // t, lambda of stage k += pend_alpha * (dt, dlam) of the last step (records and steps of the stage are staged or
    // read from the workspace); the updated records go back to the workspace for the later passes
    R pend_alpha = R(0);
    __device__ __forceinline__ void apply_pending_update(int k) {
        if (pend_alpha == R(0)) return;
        const R a = pend_alpha;
        for (int r = lane; r < NROW(); r += kTS) {
            QuadR q = recs_tl(k)[r];
            const QuadF dd = recs_dd(k)[r];
#pragma unroll
            for (int c = 0; c < 4; ++c) q.v[c] += a * R(dd.v[c]);
            if constexpr (kStageTT) reinterpret_cast<QuadR*>(sTL)[r] = q;
            *side_tl(k, r) = q;
        }
        tsync();
    }
    __device__ bool pass_factor_predict() {
        constexpr int nq = D::nq, nx = D::nx;
        const int nu = NU(), nz = NZ();
        bool ok = true;
        for (int i = lane; i < nx; i += kTS) sPv[i] = F(0);
        tsync();
        pass_fence();
        if constexpr (kStageFB) {
            tt_issue(NN(), true);
            sm_issue(NN(), false, true);
            if constexpr (kStageTT) cp_commit();            // (the wide-stage kernels only issue L2 prefetches here)
        }
        for (int k = NN(); k >= 0; --k) {
            long long f0 = UB_CLK();
            if constexpr (kStageTT) cp_wait<0>();           // side records, vectors and C rows of stage k are staged
            else if (k < NN() && NEQ() > 0) load_C(k);
            apply_pending_update(k);
            stage_gradient(k, false, F(0));
            R* gp = GPk(k);
            for (int i = lane; i < nz; i += kTS) gp[i] = sVec[i];
            if (k < NN()) {
                R* ve = wsr<R>(oVE()) + k * NEQ();
                for (int i = lane; i < NEQ(); i += kTS) ve[i] = sFv[i];
            }
            long long f1 = UB_CLK();
            UB_ACC(t_g, f1 - f0);
            // the forces go first: their elimination uses the (still idle) stage-matrix buffer as scratch
            bool fok = true;
            if (k < NN() && NFC() > 0) {
                force_block_D(k);
                fok = force_block_factor(k);
                if (!fok && nan_reason == 0) nan_reason = 6;
            }
            long long f2 = UB_CLK();
            UB_ACC(t_f2, f2 - f1);
            if (k < NN()) assign_dynamics_hessian();
            build_stage_matrix(k, k < NN());
            if constexpr (kStageFB) {                       // next stage's records / vectors / C rows arrive during the factorisation
                tsync();
                tt_issue(k - 1, true);
                sm_issue(k - 1, false, true);
                c_issue(k - 1);
                if constexpr (kStageTT) cp_commit();
            }
            long long f3 = UB_CLK();
            UB_ACC(t_f1, f3 - f2);
            if (k < NN()) {
                // reduced right-hand side [m_j; m_x]; G' g_lambda joins in the Schur update
                for (int i = lane; i < nq; i += kTS) sRv[i] = F(sVec[i]);
                for (int j = lane; j < nx; j += kTS) sRv[nq + j] = F(sVec[nu + j]);
                tsync();
                ok &= fok;
                add_dynamics_gradient(sRv);                  // uses p_{k+1} in sPv
                F* Fk = ws + oFAC() + k * FSTRIDE();
                // factor, forward substitution (w in sRv[0, nq)), p -> sPv, P -> sP, factor block -> workspace
                ok &= stage_factor_blocked(sRv, Fk, NFC() > 0 ? NEQ() : 0);
                UB_ACC(t_f3, UB_CLK() - f3);
                F* Wk = ws + oWF() + k * nq;
                for (int j = lane; j < nq; j += kTS) Wk[j] = sRv[j];
            } else {
                for (int i = lane; i < nx; i += kTS) sPv[i] = F(sVec[nu + i]);
                copy_cost_to_go();
            }
            tsync();
        }
        if constexpr (kStageTT) cp_wait<0>();               // retire the (empty) group issued at k = 0
        return ok;
    }

    // Pass C (backward, corrector): gradient = stored predictor gradient + the side terms that change
    // with the centring target and the second-order correction; backward vector step with stored factors.
    __device__ void pass_backward_corrector(F target_mu) {
        constexpr int nq = D::nq, nx = D::nx;
        const int nu = NU(), nz = NZ();
        for (int i = lane; i < nx; i += kTS) sPv[i] = F(0);
        tsync();
        const int nbx = NBOXU() + nx;
        pass_fence();
        // cp.async group schedule: [records, gradient, force bundle](k) is committed before FAC(k); every wait leaves
        // exactly one younger group in flight (none at the terminal stage)
        if constexpr (kStageFB) {
            tt_issue(NN(), true);
            sm_issue(NN(), true, false);
            cp_commit();
        }
        for (int k = NN(); k >= 0; --k) {
            if constexpr (kStageFB) {
                if (k == NN()) cp_wait<0>();
                else cp_wait<1>();
            } else {
                load_force_block(k);
            }
            const R* gpk = st_gp(k);
            for (int i = lane; i < nz; i += kTS) sVec[i] = gpk[i];
            tsync();
            // (corr - target) / (t + eps lam) per side: small terms, F arithmetic
            for (int r = lane; r < nbx; r += kTS) {
                const int fam = r < NBOXU() ? 0 : 1;
                if (!row_valid(k, fam)) continue;
                const int m = fam == 0 ? r : nu + (r - NBOXU());
                const F eps = F(row_eps(fam));
                const QuadR q = recs_tl(k)[r];
                const QuadF dd = recs_dd(k)[r];
                const F l0 = F(q.v[2]), l1 = F(q.v[3]);
                sVec[m] += R(fdiv(dd.v[0] * dd.v[2] - target_mu, F(q.v[0]) + eps * l0) -
                             fdiv(dd.v[1] * dd.v[3] - target_mu, F(q.v[1]) + eps * l1));
            }
            tsync();
            if (NFRIC() > 0 && k < NN()) {
                const F eps = F(row_eps(2));
                for (int c = lane; c < NC(); c += kTS) {
                    R g0 = 0, g1 = 0, g2 = 0;
                    for (int which = 0; which < 5; ++which) {
                        const int r = nbx + 5 * c + which;
                        const V3<R> a = fric_coeff(c, which);
                        const QuadR q = recs_tl(k)[r];
                        const QuadF dd = recs_dd(k)[r];
                        const R cf = R(fdiv(dd.v[0] * dd.v[2] - target_mu, F(q.v[0]) + eps * F(q.v[2])));
                        g0 += cf * a.x;
                        g1 += cf * a.y;
                        g2 += cf * a.z;
                    }
                    sVec[nq + 3 * c] += g0;
                    sVec[nq + 3 * c + 1] += g1;
                    sVec[nq + 3 * c + 2] += g2;
                }
                tsync();
            }
            if (NOBS() > 0 && k >= 1 && k < NN()) {
                const F eps = F(row_eps(3));
                R* crow = sScr;
                for (int i = lane; i < NOBS(); i += kTS) {
                    const int r = nbx + NFRIC() + i;
                    const QuadR q = recs_tl(k)[r];
                    const QuadF dd = recs_dd(k)[r];
                    crow[i] = R(fdiv(dd.v[0] * dd.v[2] - target_mu, F(q.v[0]) + eps * F(q.v[2])));
                }
                tsync();
                if (lane < OBSW()) {
                    R acc = 0;
                    for (int i = 0; i < NOBS(); ++i) acc += crow[i] * R(ws[oLJO() + (k * NOBS() + i) * OBSW() + lane]);
                    sVec[nu + lane] += acc;
                }
                tsync();
            }
            if (k == NN()) {
                for (int i = lane; i < nx; i += kTS) sPv[i] = F(sVec[nu + i]);
                if constexpr (kStageFB) {
                    tsync();
                    tt_issue(k - 1, true);
                    sm_issue(k - 1, true, false);
                    fb_issue(k - 1);
                    cp_commit();
                    fac_issue(k - 1, (k - 1) & 1);
                    cp_commit();
                }
                tsync();
                continue;
            }
            if (NFC() > 0) {
                // force block with the new gradient: q = D^-1 m_f, g_lambda = L^-1 (v - Df q); both replace the
                // predictor's in the stored bundle
                force_q(sD, sVec + nq, sFq);
                tsync();
                const R* ve = wsr<R>(oVE()) + k * NEQ();
                for (int r = lane; r < NEQ(); r += kTS) sFv[r] = ve[r] - eq_force_dot(r, sFq);
                tsync();
                force_Linv_mul(sS, sFv, sFg);
                tsync();
                R* gl = reinterpret_cast<R*>(bundle(k) + obGl());
                for (int r = lane; r < NEQ(); r += kTS) gl[r] = sFg[r];
                R* qf = reinterpret_cast<R*>(bundle(k) + obQ());
                for (int i = lane; i < NFC(); i += kTS) qf[i] = sFq[i];
                reduced_rhs(k, sC, nx, sFg);
            } else {
                reduced_rhs(k, nullptr, 0, nullptr);
            }
            if constexpr (kStageFB) {
                tsync();
                tt_issue(k - 1, true);
                sm_issue(k - 1, true, false);
                fb_issue(k - 1);
                cp_commit();
            }
            add_dynamics_gradient(sRv);
            const F* Fb = sM;
            if constexpr (kStageFB) {
                cp_wait<1>();
                fac_issue(k - 1, (k - 1) & 1);
                cp_commit();
                Fb = sM + (k & 1) * FSTRIDE();
            } else {
                const F* __restrict__ src = ws + oFAC() + k * FSTRIDE();
                for (int i = lane; i < FSTRIDE(); i += kTS) sM[i] = src[i];
                tsync();
            }
            F* Wk = ws + oWF() + k * nq;
            for (int j = 0; j < nq; ++j) {
                const F wj = sRv[j] * Fb[fidx(j, j)];
                tsync();
                if (lane == 0) sRv[j] = wj;
                for (int i = j + 1 + lane; i < nq; i += kTS) sRv[i] -= Fb[fidx(i, j)] * wj;
                tsync();
            }
            for (int j = lane; j < nq; j += kTS) Wk[j] = sRv[j];
            for (int i = lane; i < nx; i += kTS) {
                F acc = sRv[nq + i];
                for (int j = 0; j < nq; ++j) acc -= Fb[fidx(nq + i, j)] * sRv[j];
                sPv[i] = acc;
            }
            tsync();
        }
        if constexpr (kStageFB) cp_wait<0>();
    }

    // Side steps of the rows of ONE stage for the stage direction d = [du; dx] (shared memory):
    // d lambda, d t per side, and the running maximum feasible step.  Directions: F arithmetic on the F images of
    // t and lambda; only the slack residual rd = d + eps lam - t is formed in double.  `rnd`: bound on what the
    // F arithmetic leaves of rd after a full step.
    __device__ __forceinline__ void stage_side_steps(int k, const F* d, bool corrector, F target_mu, F& amax, F& rnd, R& s1, R& s2) {
        const R* zk = st_z(k);
        const F cm = corrector ? F(1) : F(0);
        constexpr F kEps = std::is_same<F, R>::value ? F(4.5e-16) : F(2.4e-7);
        for (int r = lane; r < NROW(); r += kTS) {
            const int fam = row_family(r);
            if (!row_valid(k, fam)) continue;
            R lb, ub;
            const R val = row_value(k, r, fam, zk, st_x(k), st_u(k), &lb, &ub);
            const F adz = F(row_dot(k, r, fam, d));
            const R eps = row_eps(fam);
            const F epsf = F(eps);
            const QuadR q = recs_tl(k)[r];
            QuadF dd = recs_dd(k)[r];
            const int nsd = fam >= 2 ? 1 : 2;
            for (int sd = 0; sd < nsd; ++sd) {
                const R t = q.v[sd], lam = q.v[2 + sd];
                const F sg = sd == 0 ? F(1) : F(-1);
                const R dist = sd == 0 ? val - lb : ub - val;
                const F rd = F(dist + eps * lam - t);
                const F tf = F(t), lf = F(lam);
                const F rc = tf * lf - target_mu + cm * dd.v[sd] * dd.v[2 + sd];
                const F iden = fdiv(F(1), tf + epsf * lf);
                const F dl = -(rc + lf * rd) * iden - (lf * iden) * sg * adz;
                const F dtt = sg * adz + epsf * dl + rd;
                dd.v[sd] = dtt;
                dd.v[2 + sd] = dl;
                // sum (t + a dt)(lam + a dlam) = sum t lam + a s1 + a^2 s2: the mean complementarity after a step of any
                // length without another pass over the records
                s1 += t * R(dl) + lam * R(dtt);
                s2 += R(dtt) * R(dl);
                rnd = max(rnd, kEps * (fabs(adz) + fabs(rd) + epsf * fabs(dl)));
                if (!(fabs(dtt) + fabs(dl) < tinf<F>())) rnd = tinf<F>();   // a step outside the range of F: not applied
                if (dtt < F(0)) amax = min(amax, fdiv(-tf, dtt));
                if (dl < F(0)) amax = min(amax, fdiv(-lf, dl));
            }
            *side_dd(k, r) = dd;
        }
    }

    // Passes B / D (forward): direction from the stored factors and w, written to DZ, with the force steps and the
    // side steps of every stage fused in.  Returns the largest feasible step in (0, 1].
    __device__ R pass_forward(bool corrector, F target_mu, R* rnd_out, R* s1_out, R* s2_out, R* smax_out) {
        constexpr int nq = D::nq, nx = D::nx;
        const int nu = NU(), nz = NZ();
        F* dxn = sDxn;          // [nx] next state direction
        F* dst = sDst;          // [nz] stage direction [dj; df; dx]
        F* dj = dst;
        F* dx = dst + nu;
        F amax = F(1), rnd = F(0);
        R s1 = R(0), s2 = R(0);
        F smax = F(0);   // largest entry of the direction (the step must be finite before it is applied)
        for (int i = lane; i < nz; i += kTS) dst[i] = F(0);
        tsync();
        pass_fence();
        // cp.async group schedule: FAC(k) is committed before [records, vectors, force bundle](k); every wait leaves
        // exactly one younger group in flight (none for the records of the terminal stage)
        if constexpr (kStageFB) {
            fac_issue(0, 0);
            w_issue(0);
            cp_commit();
            tt_issue(0, true);
            sm_issue(0, false, false);
            fb_issue(0);
            cp_commit();
        }
        for (int k = 0; k <= NN(); ++k) {
            if (k < NN()) {
                const F* Fb = sM;
                const F* Wk = ws + oWF() + k * nq;
                F wreg = F(0);
                if constexpr (kStageFB) {
                    cp_wait<1>();
                    // w_k leaves its (single) staging slot before w_{k+1} is requested
                    if (lane < nq) wreg = sSmW[lane];
                    tsync();
                    fac_issue(k + 1, (k + 1) & 1);
                    w_issue(k + 1);
                    cp_commit();
                    Fb = sM + (k & 1) * FSTRIDE();
                } else {
                    const F* __restrict__ src = ws + oFAC() + k * FSTRIDE();
                    for (int i = lane; i < FSTRIDE(); i += kTS) sM[i] = src[i];
                    if (lane < nq) wreg = Wk[lane];
                    tsync();
                }
                // s = w + Y' dx
                if (lane < nq) {
                    F acc = wreg;
                    for (int i = 0; i < nx; ++i) acc += Fb[fidx(nq + i, lane)] * dx[i];
                    dj[lane] = acc;
                }
                tsync();
                for (int j = nq - 1; j >= 0; --j) {
                    const F uj = -dj[j] * Fb[fidx(j, j)];
                    tsync();
                    for (int i = lane; i < j; i += kTS) dj[i] += Fb[fidx(j, i)] * uj;
                    if (lane == 0) dj[j] = uj;
                    tsync();
                }
                if constexpr (kStageFB) cp_wait<1>();       // records, vectors and the force bundle of stage k
                else load_force_block(k);
                if (NFC() > 0) force_block_step();
            } else {
                if constexpr (kStageFB) cp_wait<0>();
                for (int j = lane; j < nu; j += kTS) dst[j] = F(0);
                tsync();
            }
            F* Dk = DZk(k);
            for (int i = lane; i < nz; i += kTS) {
                Dk[i] = dst[i];
                smax = max(smax, fabs(dst[i]));
            }
            stage_side_steps(k, dst, corrector, target_mu, amax, rnd, s1, s2);
            if constexpr (kStageFB) {
                tsync();
                tt_issue(k + 1, true);
                sm_issue(k + 1, false, false);
                fb_issue(k + 1);
                cp_commit();
            }
            if (k < NN()) {
                if (lane < nq) {
                    const F dt = C.dt;
                    const F q = dx[lane], v = dx[nq + lane], a = dx[2 * nq + lane], j = dj[lane];
                    dxn[lane] = q + dt * v + F(0.5) * dt * dt * a + dt * dt * dt / F(6) * j;
                    dxn[nq + lane] = v + dt * a + F(0.5) * dt * dt * j;
                    dxn[2 * nq + lane] = a + dt * j;
                }
                tsync();
                for (int i = lane; i < nx; i += kTS) dx[i] = dxn[i];
                tsync();
            }
        }
        if constexpr (kStageFB) cp_wait<0>();
        *rnd_out = R(tmax(rnd));
        *s1_out = tsum(s1);
        *s2_out = tsum(s2);
        *smax_out = R(tmax(smax));
        return R(tmin(amax));
    }

    // Interior-point QP solve around the current (X, U).  Leaves the step in Z
    // and the factors of the last iteration in FAC.  Returns iterations used;
    // *converged, *decr as in orc::solve_qp_ipm.  A breakdown of the F-precision factorisation (possible only on
    // problems the iteration does not converge on: infeasible hard constraints) ends the iteration at the last
    // finite iterate, reported like the iteration cap.
    __device__ int solve_qp(bool* converged, R* decr, bool* finite) {
        constexpr int nq = D::nq, nx = D::nx;
        const int nu = NU(), nz = NZ(), N = NN();
        *converged = false;
        *finite = true;
        const long long c_qi = UB_CLK();
        // dynamics-feasible start: du = 0, dx_0 = 0, dx_{k+1} = A dx_k + gap_k
        R* Zall = wsr<R>(oZ());
        for (int idx = lane; idx < (N + 1) * nz; idx += kTS) Zall[idx] = R(0);
        tsync();
        if (lane < nq) {
            const R dt = PR.dt;
            R q = 0, v = 0, a = 0;
            for (int k = 0; k < N; ++k) {
                const R* gap = wsr<R>(oGAP()) + k * nx;
                const R qn = q + dt * v + R(0.5) * dt * dt * a + gap[lane];
                const R vn = v + dt * a + gap[nq + lane];
                const R an = a + gap[2 * nq + lane];
                q = qn; v = vn; a = an;
                R* zn = Zk(k + 1);
                zn[nu + lane] = q;
                zn[nu + nq + lane] = v;
                zn[nu + 2 * nq + lane] = a;
            }
        }
        tsync();
        init_eq_weights();
        // slack / multiplier initialisation, with the residual summary of the starting point (mean complementarity and
        // the largest slack residual) collected on the way
        int nsides_l = 0;
        R m_l = 0, rd_l = 0;
        for (int k = 0; k <= N; ++k) {
            const R* zk = Zk(k);
            for (int r = lane; r < NROW(); r += kTS) {
                const int fam = row_family(r);
                QuadR q;
                QuadF dd;
                q.v[0] = q.v[1] = R(1);
                q.v[2] = q.v[3] = R(0);
                dd.v[0] = dd.v[1] = dd.v[2] = dd.v[3] = F(0);
                if (row_valid(k, fam)) {
                    R lb, ub;
                    const R val = row_value(k, r, fam, zk, X + k * NX(), U + k * NU(), &lb, &ub);
                    const R eps = row_eps(fam);
                    q.v[0] = max(val - lb, PR.thr0);
                    q.v[2] = PR.mu0 / q.v[0];
                    rd_l = max(rd_l, fabs(val - lb + eps * q.v[2] - q.v[0]));
                    m_l += q.v[0] * q.v[2];
                    ++nsides_l;
                    if (fam < 2) {
                        q.v[1] = max(ub - val, PR.thr0);
                        q.v[3] = PR.mu0 / q.v[1];
                        rd_l = max(rd_l, fabs(ub - val + eps * q.v[3] - q.v[1]));
                        m_l += q.v[1] * q.v[3];
                        ++nsides_l;
                    }
                }
                *side_tl(k, r) = q;
                *side_dd(k, r) = dd;
            }
        }
        const int nsides = int(tsum(R(nsides_l)) + R(0.5));
        tsync();
        R last_alpha = R(0), last_step = tinf<R>();
        pend_alpha = R(0);
        int iters = 0;
        R mu = nsides > 0 ? tsum(m_l) / R(nsides) : R(0), rdmax = tmax(rd_l);
        R pinf = R(0);
        // hard equality rows: multiplier update y += rho e (update = true) and their largest residual
        auto eq_pass = [&](bool update) {
            R pv = R(0);
            if (!C.soft_poly) {
                if (NEQ() > 0 && NEQ() <= 8) {
                    // one body: four lanes share a row's dot product and read the C rows straight from the workspace
                    // (no staging copy, no barrier: the loads of all stages are independent)
                    const int ne = NEQ(), i = lane >> 2, part = lane & 3;
                    for (int k = 0; k < N; ++k) {
                        const R* zk = Zk(k);
                        R v = R(0);
                        if (i < ne) {
                            const F* c = ws + oLC() + (k * ne + i) * nx;
                            for (int j = part; j < nx; j += 4) v += R(c[j]) * zk[nu + j];
                            if (part == 0) v += wsr<R>(oLG())[k * ne + i] + eq_force_dot(i, zk + nq);
                        }
                        v += __shfl_xor_sync(FULL, v, 1);
                        v += __shfl_xor_sync(FULL, v, 2);
                        if (i < ne && part == 0) {
                            const R rho = rho_eq(k)[i];
                            if (rho > R(0)) {
                                if (update) y_eq(k)[i] += rho * v;
                                pv = max(pv, fabs(v));
                            }
                        }
                    }
                    tsync();
                } else
                for (int k = 0; k < N; ++k) {
                    if (NEQ() == 0) break;
                    load_C(k);
                    const R* zk = Zk(k);
                    for (int i = lane; i < NEQ(); i += kTS) {
                        const R rho = rho_eq(k)[i];
                        if (rho > R(0)) {
                            const R e = eq_value_stage(k, i, zk);
                            if (update) y_eq(k)[i] += rho * e;
                            pv = max(pv, fabs(e));
                        }
                    }
                    tsync();
                }
                const R* zk = Zk(N);
                for (int i = lane; i < NTERM(); i += kTS) {
                    const R rho = rho_eq(N)[i];
                    if (rho > R(0)) {
                        const R e = eq_value_term(i, zk);
                        if (update) y_eq(N)[i] += rho * e;
                        pv = max(pv, fabs(e));
                    }
                }
                tsync();
            }
            return tmax(pv);
        };
        pinf = eq_pass(false);
        UB_ACC(t_qi, UB_CLK() - c_qi);
        R pinf_prev = tinf<R>();
        int stall = 0;
        for (int it = 0; it < C.qp_iter_max; ++it) {
            const long long c_al = UB_CLK();
            align_wait();   // meet the other warps of the CTA: same code at the same time (see CtaAlign)
            const long long c_it = UB_CLK();
            UB_ACC(t_al, c_it - c_al);
            if (it > 0 && mu <= R(2) * PR.mu_target && rdmax <= PR.qp_tol && last_alpha >= R(0.5) &&
                (pinf <= PR.qp_tol || last_step <= PR.qp_tol)) {
                *converged = true;
                break;
            }
            // hopeless hard equality rows (orc::solve_qp_ipm): at the observed linear rate they cannot reach the
            // tolerance before the iteration cap, three iterations in a row -> the QP ends as not converged
            {
                bool hopeless = false;
                if (!C.soft_poly && it >= 4 && pinf > PR.qp_tol && pinf_prev < tinf<R>())
                    hopeless = hopeless_rate(pinf, pinf_prev, PR.qp_tol, it, C.qp_iter_max);
                stall = hopeless ? stall + 1 : 0;
                pinf_prev = pinf;
                if (stall >= 3) break;
            }
            if (!pass_factor_predict()) {
                if constexpr (std::is_same<F, R>::value) {   // validation kernels: as the oracle, a failed factorisation ends the solve
                    *finite = false;
                    iters = it + 1;
                    if (nan_reason == 0) nan_reason = 1;
                }
                break;
            }
            iters = it + 1;
            long long c2 = UB_CLK();
            UB_ACC(t_fac, c2 - c_it);
            R target_mu = PR.mu_target;
            R alpha = R(1);
            // predictor (pass 0) and corrector (pass 1) share ONE inlined copy of the forward pass
            R a_fwd = R(1), rnd_gap = R(0), cs1 = R(0), cs2 = R(0), dzmax = R(0);
#pragma unroll 1
            for (int pass = 0; pass < (nsides > 0 ? 2 : 1); ++pass) {
                if (pass == 1) {
                    // mean complementarity after the affine step -> centring target (Mehrotra)
                    const R a_aff = a_fwd;
                    const R mu_aff = mu + (a_aff * cs1 + a_aff * a_aff * cs2) / R(nsides);
                    const R ratio = mu_aff / mu;
                    target_mu = max(ratio * ratio * ratio * mu, PR.mu_target);
                    long long c3 = UB_CLK();
                    UB_ACC(t_swp, c3 - c2);
                    c2 = c3;
                    pass_backward_corrector(F(target_mu));
                }
                a_fwd = pass_forward(pass == 1, pass == 1 ? F(target_mu) : F(0), &rnd_gap, &cs1, &cs2, &dzmax);
            }
            if (nsides > 0) alpha = min(R(1), R(0.995) * a_fwd);
            UB_ACC(t_side, UB_CLK() - c2);
            // the step must be finite before it is applied
            const R stepmax = fabs(alpha * dzmax);   // = max |alpha dz| (collected by the forward pass)
            if (!(stepmax < tinf<R>()) || !(alpha > R(0)) || !(alpha <= R(1)) || !(rnd_gap < tinf<R>())) {
                if constexpr (std::is_same<F, R>::value) {
                    *finite = false;
                    if (nan_reason == 0) nan_reason = !(alpha <= R(1)) ? 5 : 2;
                }
                break;
            }
            // update z; t and lambda follow lazily, stage by stage, when the next factor pass visits them
            // (apply_pending_update), and their new mean complementarity comes from the sums of the forward pass
            for (int idx = lane; idx < (N + 1) * nz; idx += kTS) Zall[idx] += alpha * R(ws[oDZ() + idx]);
            pend_alpha = alpha;
            tsync();
            mu = nsides > 0 ? mu + (alpha * cs1 + alpha * alpha * cs2) / R(nsides) : R(0);
            // the rows are linear: the slack residual contracts by (1 - alpha), up to the rounding of the stored steps
            // (zero in the fp64 kernels, where this is the exact recursion)
            rdmax = (R(1) - alpha) * rdmax + alpha * rnd_gap;
            if (!C.soft_poly) pinf = eq_pass(true);
            last_alpha = alpha;
            last_step = stepmax;
            if (!(mu < tinf<R>())) {
                *finite = false;
                if (nan_reason == 0) nan_reason = 4;
                break;
            }
        }
        *decr = last_step;
        return iters;
    }

    // Feedback gains K_k = -Huu^{-1} Hux from the stored factors (optional output): jerk rows from the Riccati
    // blocks, force rows -D^-1 Df' L^-T G from the force block.
    __device__ void write_gains(F* Kout, int stages) {
        constexpr int nq = D::nq, nx = D::nx;
        const int nu = NU();
        F* Fb = sM;
        F* col = sDst;
        for (int k = 0; k < stages; ++k) {
            const F* Fg = ws + oFAC() + k * FSTRIDE();
            for (int idx = lane; idx < FSTRIDE(); idx += kTS) Fb[idx] = Fg[idx];
            tsync();
            load_force_block(k);
            for (int xcol = 0; xcol < nx; ++xcol) {
                for (int j = lane; j < nq; j += kTS) col[j] = Fb[fidx(nq + xcol, j)];
                tsync();
                for (int j = nq - 1; j >= 0; --j) {
                    const F uj = -col[j] * Fb[fidx(j, j)];
                    tsync();
                    for (int i = lane; i < j; i += kTS) col[i] += Fb[fidx(j, i)] * uj;
                    if (lane == 0) col[j] = uj;
                    tsync();
                }
                for (int j = lane; j < nq; j += kTS) Kout[(k * nu + j) * nx + xcol] = col[j];
                tsync();
                if (NFC() > 0) {
                    // unit state step along xcol: lambda = L^-T G e, df = -D^-1 Df' lambda (no constant terms)
                    const int ne = NEQ(), nf = NF();
                    for (int r = lane; r < ne; r += kTS) sFv[r] = R(sC[r * nx + xcol]);
                    tsync();
                    force_LinvT_mul(sS, sFv, sFl);
                    tsync();
                    const R* Dg = sD;
                    const F* Dc = sDFC;
                    for (int c = lane; c < NC(); c += kTS) {
                        R u[3] = {0, 0, 0};
                        for (int side = 0; side < 2; ++side) {
                            const int b = side == 0 ? P.cb2[c] : P.cb1[c];
                            if (b < 0) continue;
                            const F* blk = Dc + ((2 * c + side) * 6) * nf;
                            for (int rr = 0; rr < 6; ++rr)
                                for (int i = 0; i < nf; ++i) u[i] += R(blk[rr * nf + i]) * sFl[6 * b + rr];
                        }
                        if (nf == 1) Kout[(k * nu + nq + c) * nx + xcol] = F(-Dg[c] * u[0]);
                        else {
                            const R* o = Dg + 9 * c;
                            for (int i = 0; i < 3; ++i)
                                Kout[(k * nu + nq + 3 * c + i) * nx + xcol] = F(-(o[3 * i] * u[0] + o[3 * i + 1] * u[1] + o[3 * i + 2] * u[2]));
                        }
                    }
                    tsync();
                }
            }
        }
    }

    // --------------------------------------------------------------- solve
    __device__ void run(const BatchArgs<F>& A, int b) {
        constexpr int nq = D::nq, nx = D::nx;
        const int nu = NU(), N = NN(), nz = NZ();
        t_lin = t_fac = t_swp = t_side = t_ls = t_f1 = t_f2 = t_f3 = t_g = t_al = t_qi = 0;
        const long long t_run0 = UB_CLK();
        long long t_init = 0;
        nan_reason = 0;
        // initial guess: DefaultInitializer = zero input, state held
        // (controller_interface.cpp:385-386); x_0 is always the observation
        {
            const int nxt = A.nxt, nxo = NXO();
            const F* x0 = A.x0 + size_t(b) * nxt;
            F* XOw = ws + oXO();
            if (!A.warm) {
                for (int idx = lane; idx < (N + 1) * nx; idx += kTS) X[idx] = x0[idx % nx];
                for (int idx = lane; idx < N * nu; idx += kTS) U[idx] = F(0);
                for (int idx = lane; idx < (N + 1) * nxo; idx += kTS) XOw[idx] = x0[nx + idx % nxo];   // state held
            } else {
                const F* Xin = A.Xin + size_t(b) * (N + 1) * nxt;
                const F* Uin = A.Uin + size_t(b) * N * nu;
                for (int idx = lane; idx < (N + 1) * nx; idx += kTS) {
                    const int k = idx / nx, i = idx % nx;
                    X[idx] = k == 0 ? x0[i] : Xin[k * nxt + i];
                }
                for (int idx = lane; idx < N * nu; idx += kTS) U[idx] = Uin[idx];
                for (int idx = lane; idx < (N + 1) * nxo; idx += kTS) {
                    const int k = idx / nxo, i = idx % nxo;
                    XOw[idx] = k == 0 ? x0[nx + i] : Xin[k * nxt + nx + i];
                }
            }
            const F* tg = A.target + size_t(b) * (N + 1) * A.tstride;
            F* tgl = ws + oTG();
            for (int idx = lane; idx < (N + 1) * 3; idx += kTS) tgl[idx] = tg[(idx / 3) * A.tstride + idx % 3];
            if (ORI())
                for (int idx = lane; idx < (N + 1) * 4; idx += kTS) ws[oTGQ() + idx] = tg[(idx / 4) * A.tstride + 3 + idx % 4];
            const F* bd = A.body ? A.body + size_t(b) * NB() * UB_BODY_PARAMS : &P.body[0][0];
            F* bdl = ws + oBD();
            for (int idx = lane; idx < NB() * UB_BODY_PARAMS; idx += kTS) bdl[idx] = bd[idx];
        }
        tsync();
        if (NEQ() > 0) build_Df();
        Perf<F> base = performance(X, U);
        UB_ACC(t_init, UB_CLK() - t_run0);
        int status = UB_STATUS_CONVERGED, qp_iters = 0, sqp_done = 0;
        F alpha = 0;
        R qp_res = 0;
        F* Xn = ws + oXN();
        F* Un = ws + oUN();
        const R* Zall = wsr<R>(oZ());
        for (int it = 0; it < max(1, C.sqp_iters); ++it) {
            ++sqp_done;
            long long c0 = UB_CLK();
            if (NXO() > 0) {
                // Newton step of the obstacle states = exact constant-acceleration rollout of the observation - iterate
                const F* o0 = ws + oXO();
                for (int idx = lane; idx < (N + 1) * P.ndyn * 3; idx += kTS) {
                    const int k = idx / (P.ndyn * 3), j = (idx / 3) % P.ndyn, c = idx % 3;
                    const F t = C.dt * F(k);
                    const F p0 = o0[j * 9 + c], v0 = o0[j * 9 + 3 + c], a0 = o0[j * 9 + 6 + c];
                    const int o = (k * P.ndyn + j) * 9;
                    ws[oDXO() + o + c] = p0 + t * v0 + F(0.5) * t * t * a0 - ws[oXO() + o + c];
                    ws[oDXO() + o + 3 + c] = v0 + t * a0 - ws[oXO() + o + 3 + c];
                    ws[oDXO() + o + 6 + c] = a0 - ws[oXO() + o + 6 + c];
                }
                tsync();
            }
            linearize();
            UB_ACC(t_lin, UB_CLK() - c0);
            if (A.stop_after == 1) break;
            bool conv, fin;
            qp_iters += solve_qp(&conv, &qp_res, &fin);
            if (A.stop_after == 2) break;
            if (!fin) {
                status = UB_STATUS_NAN;
                break;
            }
            if (!conv) status = UB_STATUS_QP_MAXITER;
            // Armijo descent metric: cost gradient (at z = 0) along the step
            F desc = 0;
            for (int k = 0; k < N; ++k) {
                const R* zk = Zall + k * nz;
                const F* x = X + k * nx;
                const F* u = U + k * nu;
                const F* Jp = ws + oLJP() + k * 3 * nq;
                for (int i = lane; i < nz; i += kTS) {
                    F g;
                    if (i < nq) g = C.dt * P.Rd[i] * u[i];
                    else if (i < nu) g = C.dt * C.fw * u[i];
                    else {
                        const int xi = i - nu;
                        g = C.dt * P.Qd[xi] * (x[xi] - P.xd[xi]);
                        if (xi < nq)
                            for (int c = 0; c < 3; ++c)
                                g += C.dt * C.Wd[c] * Jp[c * nq + xi] * (F(wsr<R>(oLR())[3 * k + c]) - target[3 * k + c]);
                        if (ORI() && xi < nq)
                            for (int c = 0; c < 3; ++c)
                                g += C.dt * P.Wo[c] * ws[oLJQ() + (k * 3 + c) * nq + xi] * F(wsr<R>(oLRO())[3 * k + c]);
                        if (IALIGN())
                            for (int r = 0; r < 2; ++r)
                                g += C.dt * P.ia_w * ws[oLJA() + (2 * k + r) * nx + xi] * ws[oLIA() + 2 * k + r];
                    }
                    desc += g * F(zk[i]);
                }
            }
            desc = tsum(desc);
            const long long c_ls = UB_CLK();
            // filter line search (ocs2 FilterLinesearch [EXT])
            const F vb = base.violation();
            bool accepted = false;
            Perf<F> pn = base;
            alpha = F(1);
            while (alpha >= C.alpha_min) {
                for (int idx = lane; idx < (N + 1) * nx; idx += kTS) {
                    const int k = idx / nx, i = idx % nx;
                    Xn[idx] = F(R(X[idx]) + R(alpha) * Zall[k * nz + nu + i]);
                }
                for (int idx = lane; idx < N * nu; idx += kTS) {
                    const int k = idx / nu, i = idx % nu;
                    Un[idx] = F(R(U[idx]) + R(alpha) * Zall[k * nz + i]);
                }
                tsync();
                pn = performance(Xn, Un, alpha);
                const F vn = pn.violation();
                if (vn > C.g_max) accepted = false;
                else if (vn < C.g_min) {
                    if (vb < C.g_min && desc < F(0)) accepted = pn.cost < base.cost + C.armijo * alpha * desc;
                    else accepted = true;
                } else {
                    accepted = (vn < (F(1) - C.gamma_c) * vb) || (pn.cost < base.cost - C.gamma_c * vb);
                }
                if (accepted) break;
                alpha *= C.alpha_decay;
            }
            UB_ACC(t_ls, UB_CLK() - c_ls);
            if (!accepted) {
                status = UB_STATUS_LS_FAILED;
                alpha = F(0);
                break;
            }
            F dxn = 0, dun = 0;
            for (int idx = lane; idx < (N + 1) * nx; idx += kTS) {
                const F d = Xn[idx] - X[idx];
                dxn += d * d;
                X[idx] = Xn[idx];
            }
            for (int idx = lane; idx < N * nu; idx += kTS) {
                const F d = Un[idx] - U[idx];
                dun += d * d;
                U[idx] = Un[idx];
            }
            for (int idx = lane; idx < (N + 1) * NXO(); idx += kTS) ws[oXO() + idx] += alpha * ws[oDXO() + idx];
            dxn = sqrt(tsum(dxn));
            dun = sqrt(tsum(dun));
            tsync();
            const F dcost = fabs(pn.cost - base.cost);
            base = pn;
            if ((dxn < C.delta_tol && dun < C.delta_tol) || (dcost < C.cost_tol && base.violation() < C.g_min)) break;
        }
        if (A.K != nullptr && A.stop_after == 0 && status != UB_STATUS_NAN)
            write_gains(A.K + size_t(b) * A.gain_stages * nu * nx, A.gain_stages);
        // solution out + NaN guard
        F bad = 0;
        {
            const int nxt = A.nxt, nxo = NXO();
            F* Xo = A.X + size_t(b) * (N + 1) * nxt;
            F* Uo = A.U + size_t(b) * N * nu;
            for (int idx = lane; idx < (N + 1) * nx; idx += kTS) {
                const F v = X[idx];
                bad += isfinite(v) ? F(0) : F(1);
                Xo[(idx / nx) * nxt + idx % nx] = v;
            }
            for (int idx = lane; idx < (N + 1) * nxo; idx += kTS) Xo[(idx / nxo) * nxt + nx + idx % nxo] = ws[oXO() + idx];
            for (int idx = lane; idx < N * nu; idx += kTS) Uo[idx] = U[idx];
            // the same rows into the peers' gathered buffers (P2P stores over NVLink)
            for (int p = 0; p < A.ngather; ++p) {
                F* Xp = A.Xg[p] + size_t(A.gather_row + b) * (N + 1) * nxt;
                F* Up = A.Ug[p] + size_t(A.gather_row + b) * N * nu;
                for (int idx = lane; idx < (N + 1) * nx; idx += kTS) Xp[(idx / nx) * nxt + idx % nx] = X[idx];
                for (int idx = lane; idx < (N + 1) * nxo; idx += kTS) Xp[(idx / nxo) * nxt + nx + idx % nxo] = ws[oXO() + idx];
                for (int idx = lane; idx < N * nu; idx += kTS) Up[idx] = U[idx];
            }
        }
        bad = tsum(bad);
        if (bad > F(0)) {
            status = UB_STATUS_NAN;
            if (nan_reason == 0) nan_reason = 3;
        }
        // status[b] is the completion flag of the instance: the host path polls it in mapped pinned memory and
        // converts the rows of finished instances while the kernel is still solving others, so every result of
        // this warp is fenced system-wide before lane 0 publishes the status (last store below)
        __threadfence_system();
        tsync();
        if (lane == 0) {
            if (A.stats) {
                F* s = A.stats + size_t(b) * UB_STATS;
                s[0] = F(qp_iters);
                s[1] = base.cost;
                s[2] = base.violation();
                s[3] = status == UB_STATUS_NAN ? F(nan_reason) : alpha;   // NaN status: where it first appeared
                s[4] = F(qp_res);
                s[5] = base.max_eq;
                s[6] = base.min_margin;
                s[7] = F(sqp_done);
                if (A.stop_after == 8) {  // profile mode: linearisation, line search, whole solve, interior-point loop
                    s[1] = F(t_lin);
                    s[2] = F(t_ls);
                    s[3] = F(UB_CLK() - t_run0);
                    s[4] = F(t_fac + t_swp + t_side);
                    s[5] = F(t_init);
                    s[6] = F(t_al);
                    s[7] = F(t_qi);
                }
                if (A.stop_after == 9) {  // profile mode: phase cycle counters replace stats[1..7]
                    s[1] = F(t_g);
                    s[2] = F(t_f1);
                    s[3] = F(t_f2);
                    s[4] = F(t_f3);
                    s[5] = F(t_fac);
                    s[6] = F(t_swp);
                    s[7] = F(t_side);
                }
                __threadfence_system();
            }
            *reinterpret_cast<volatile int32_t*>(A.status + b) = status;
        }
    }
};

template <typename F, typename D>
__global__ void __launch_bounds__(512, 1) solve_batch_kernel(const __grid_constant__ DevProblem<F> Pc, const DevProblem<F>* __restrict__ Pg,
                                                          const DevProblem<double>* __restrict__ Pgr, Layout L, BatchArgs<F> A,
                                                          int teams_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // CTA-shared copies of the problem constants: F for the matrix arithmetic, double for residuals / linearisation
    // (one and the same for the fp64 validation kernels)
    constexpr bool kSame = std::is_same<F, double>::value;
    constexpr size_t offF = (sizeof(DevProblem<F>) + 15) / 16 * 16;
    constexpr size_t offP = offF + (kSame ? 0 : (sizeof(DevProblem<double>) + 15) / 16 * 16);
    constexpr size_t off = offP + sizeof(CtaAlign);   // (16 bytes: the launcher counts them in)
    DevProblem<F>* Ps = reinterpret_cast<DevProblem<F>*>(smem_raw);
    DevProblem<double>* Psr = kSame ? reinterpret_cast<DevProblem<double>*>(smem_raw) : reinterpret_cast<DevProblem<double>*>(smem_raw + offF);
    {
        const int nwords = sizeof(DevProblem<F>) / 4;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Pg);
        uint32_t* dst = reinterpret_cast<uint32_t*>(Ps);
        for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
        if constexpr (!kSame) {
            const int nw2 = sizeof(DevProblem<double>) / 4;
            const uint32_t* s2 = reinterpret_cast<const uint32_t*>(Pgr);
            uint32_t* d2 = reinterpret_cast<uint32_t*>(Psr);
            for (int i = threadIdx.x; i < nw2; i += blockDim.x) d2[i] = s2[i];
        }
    }
    CtaAlign* al = reinterpret_cast<CtaAlign*>(smem_raw + offP);
    const int cta_members = min(teams_per_cta, A.n_slots - int(blockIdx.x) * teams_per_cta);
    // alignment groups: A.align = warps per group (0: off); at most four groups
    int n_groups = A.align > 0 ? (cta_members + A.align - 1) / A.align : 1;
    n_groups = max(1, min(4, n_groups));
    if (threadIdx.x < 4) {
        int members = 0;
        for (int t = int(threadIdx.x); t < cta_members; t += n_groups) ++members;
        al->word[threadIdx.x] = (unsigned(members) << 12) | (unsigned(members) << 6);   // all members still working
    }
    __syncthreads();
    const int team = int(threadIdx.x) / WARP, lane = threadIdx.x % WARP;
    const int slot = blockIdx.x * teams_per_cta + team;
    if (slot >= A.n_slots) return;
    Layout Lk = L;
    if constexpr (D::kStatic) Lk = D::template layout<F>();  // compile-time offsets (host passes the same numbers)
    // The per-warp shared-memory offset and the per-slot workspace offset pass through an opaque move: the
    // compiler then keeps them in a register instead of re-deriving them from threadIdx / blockIdx at every use
    // (measured: that rematerialisation was ~8 % of all executed instructions).
    uint32_t sm_off = uint32_t(off + size_t(team) * Lk.s_total * sizeof(F));
    asm volatile("mov.u32 %0, %0;" : "+r"(sm_off));
    unsigned long long ws_off = (unsigned long long)(slot) * (unsigned long long)(Lk.total);
    asm volatile("mov.u64 %0, %0;" : "+l"(ws_off));
    F* sm = reinterpret_cast<F*>(smem_raw + sm_off);
    Solver<F, D> S(*Ps, Pc, *Psr, L, lane);
    S.ws = A.ws + ws_off;
    S.X = S.ws + Lk.XW;
    S.U = S.ws + Lk.UW;
    S.target = S.ws + Lk.TG;
    S.body = S.ws + Lk.BD;
    S.sM = sm + Lk.sM;
    S.sP = sm + Lk.sP;
    S.sPv = sm + Lk.sPv;
    S.sFB = sm + Lk.sFB;
    S.sC = S.sFB + Lk.bG;
    S.sS = reinterpret_cast<double*>(S.sFB + Lk.bL);
    S.sD = reinterpret_cast<double*>(S.sFB + Lk.bD);
    S.sFg = reinterpret_cast<double*>(S.sFB + Lk.bGl);
    S.sFq = reinterpret_cast<double*>(S.sFB + Lk.bQ);
    S.sGf = sm + Lk.sGf;
    S.sDFC = sm + Lk.sDFC;
    S.sUS = reinterpret_cast<double*>(sm + Lk.sUS);
    S.sCst = sm + Lk.sCst;
    S.cC = (D::kStatic && D::nrow <= UB_STAGE_ROWS_MAX) ? S.sCst : S.sC;
    S.sFv = reinterpret_cast<double*>(sm + Lk.sFv);
    S.sFl = reinterpret_cast<double*>(sm + Lk.sFl);
    S.sVec = reinterpret_cast<double*>(sm + Lk.sVec);
    S.sDst = sm + Lk.sDst;
    S.sDxn = sm + Lk.sDxn;
    S.sRv = sm + Lk.sRv;
    S.sScr = reinterpret_cast<double*>(sm + Lk.sScr);
    S.sTL = reinterpret_cast<double*>(sm + Lk.sTL);
    S.sDD = sm + Lk.sDD;
    S.sSmZ = reinterpret_cast<double*>(sm + Lk.sSmZ);
    S.sSmX = sm + Lk.sSmX;
    S.sSmU = sm + Lk.sSmU;
    S.sSmJ = sm + Lk.sSmJ;
    S.sSmW = sm + Lk.sSmW;
    S.bar_init(sm + Lk.sBar);
    if (A.queue == nullptr) {   // static mode (test aid)
        S.run(A, slot);
        return;
    }
    if (A.align > 0) {
        const int gid = team % n_groups;
        int members = 0;
        for (int t = gid; t < cta_members; t += n_groups) ++members;
        if (members > 1) S.al = &al->word[gid];
    }
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(A.queue, 1);
        b = __shfl_sync(FULL, b, 0);
        if (b >= A.B) break;
        S.run(A, b);
        S.tsync();
    }
    // out of work: keep attending the meetings of the CTA until every warp is
    if (S.al != nullptr) {
        if (lane == 0) atomicSub(S.al, 1u << 6);
        __syncwarp();
        while (!S.align_wait()) {
        }
    }
}

}  // namespace ub
