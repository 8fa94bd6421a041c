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
#pragma once
#include <math_constants.h>

#include "ub_device.cuh"

namespace ub {

// Per-problem workspace layout in units of T.  One constexpr function serves the host (make_layout in
// ub_api.cu) and the kernels specialised on compile-time dimensions, where every offset becomes an
// immediate of the load/store instruction.
struct Layout {
    int Z, DZ, GAP, LG, LCT, LR, LJP, LHO, LJO, DF, RHOE, YE, RHOT, YT, TT, LAM, XW, UW, sTT, FAC, WF, XN, UN;
    int total;
    // shared memory (units of T, per warp)
    int sM, sP, sPv, sSA, sV, s_total;
    int sSm;  // staged per-stage vectors [z | x | u | Jp | w] (with the staged side records)
    int sRed; // team reduction / broadcast scratch (16 values)
    // instance-local copies of the desired positions [N+1, 3] and the body parameters [nb, 10]
    int TG, BD;
    // inertial-alignment cost rows: values [N, 2] and Jacobians [N, 2, nx] (only when that cost is enabled)
    int LIA, LJA;
    // dynamic-obstacle states: iterate [N+1, nxo] and Newton step (exact rollout - iterate) [N+1, nxo]
    int XO, DXO;
};
// side records of a stage are staged in shared memory (cp.async, one stage ahead) up to this many rows
#define UB_STAGE_ROWS_MAX 64
struct LayoutDims {
    int N, nq, nx, nu, neq, nfc, nterm, nrow, nobs, nb, tsize;
    int nia = 0;   // rows of the inertial-alignment cost (0 or 2)
    int obsw = 0;  // width of the obstacle-family rows (0 -> nq)
    int nxo = 0;   // dynamic-obstacle states (9 per obstacle)
};
__host__ __device__ constexpr int ub_round4(int n) { return (n + 3) / 4 * 4; }
__host__ __device__ constexpr Layout compute_layout(const LayoutDims d) {
    Layout L{};
    const int N = d.N, nx = d.nx, nu = d.nu, nz = d.nu + d.nx, nq = d.nq;
    int o = 0;
    const int ldm = nz | 1, ldf = nu | 1;
    L.Z = o;    o += ub_round4((N + 1) * nz);
    L.DZ = o;   o += ub_round4((N + 1) * nz);
    L.GAP = o;  o += ub_round4(N * nx);
    L.LG = o;   o += ub_round4(N * d.neq);
    L.LCT = o;  o += ub_round4(N * d.neq * nz);
    L.LR = o;   o += ub_round4((N + 1) * 3);
    L.LJP = o;  o += ub_round4((N + 1) * 3 * nq);
    L.LHO = o;  o += ub_round4((N + 1) * d.nobs);
    L.LJO = o;  o += ub_round4((N + 1) * d.nobs * (d.obsw > 0 ? d.obsw : nq));
    L.DF = o;   o += ub_round4(d.neq * d.nfc);
    L.RHOE = o; o += ub_round4(N * d.neq);
    L.YE = o;   o += ub_round4(N * d.neq);
    L.RHOT = o; o += ub_round4(d.nterm);
    L.YT = o;   o += ub_round4(d.nterm);
    L.TT = o;   o += ub_round4((N + 1) * d.nrow * 8);  // interleaved side records {t, lam, dt, dlam} x {lo, hi}
    L.LAM = o;  o += ub_round4((N + 1) * nz);          // GP: predictor stage gradients kept for the corrector
    // factor block of a stage: [L; Y] row-major with leading dimension nu|1 (generic kernels) or column-major with
    // column length nz+1 (blocked kernels, UB_BLOCKED_*); sized for either
    const int fs_row = (nz * ldf + 3) & ~3, fs_col = (nu * (nz + 1) + 3) & ~3;
    const int fstride = fs_row > fs_col ? fs_row : fs_col;
    L.FAC = o;  o += ub_round4(N * fstride);
    L.WF = o;   o += ub_round4(N * nu);
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
    L.total = o;
    int s = 0;
    L.sM = s;   s += ub_round4(nz * ldm > fstride ? nz * ldm : fstride);
    L.sP = s;   s += ub_round4(nx * nx);
    L.sPv = s;  s += ub_round4(nx);
    L.sSA = s;  s += ub_round4((d.neq > 3 ? d.neq : 3) * nz);
    L.sV = s;   s += ub_round4(5 * nz + 64);  // [4 nz, ...) doubles as per-row scratch (>= max(neq, nobs, 3) entries)
    L.sTT = s;  s += (d.nrow <= UB_STAGE_ROWS_MAX) ? ub_round4(d.nrow * 8) : 0;  // staged side records of one stage
    L.sSm = s;
    s += (d.nrow <= UB_STAGE_ROWS_MAX) ? ub_round4(nz) + ub_round4(nx) + 2 * ub_round4(nu) + ub_round4(3 * nq) : 0;
    L.sRed = s;
    s += 16;
    L.s_total = s;
    return L;
}

template <typename T>
struct BatchArgs {
    const T* x0;      // [B, nx]
    const T* target;  // [B, N+1, 3]
    const T* body;    // [B, nb, 10] or null
    T* X;             // [B, N+1, nx]  solution out (device memory, or mapped pinned host memory: the host path lets
    T* U;             // [B, N, nu]    the kernel write results home while other instances are still being solved)
    const T* Xin;     // warm start in (UB_WARM_START); may alias X
    const T* Uin;
    T* K;             // [B, N, nu, nx] or null
    int32_t* status;  // [B]
    T* stats;         // [B, UB_STATS] or null
    T* ws;            // [B, layout.total]
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
};


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <typename T>
__device__ __forceinline__ T tinf() { return T(1e30); }

template <typename T>
struct Perf {
    T cost, dyn, eq, ineq, max_eq, min_margin;
    __device__ __forceinline__ T violation() const { return sqrt(dyn + eq + ineq); }
};

// Compile-time problem dimensions (specialised kernels) or run-time ones (generic kernel).
template <int NQ_, int NF_, int NC_, int NB_, int NOBS_ = 0, int N_ = 20>
struct StaticDims {
    static constexpr bool kStatic = true;
    static constexpr int nq = NQ_, nf = NF_, nc = NC_, nb = NB_, nobs = NOBS_, N = N_;
    static constexpr int nx = 3 * NQ_, nfc = NF_ * NC_, nu = NQ_ + NF_ * NC_, nz = nu + nx;
    static constexpr int neq = 6 * NB_, nfric = (NF_ == 3) ? 5 * NC_ : 0;
    static constexpr int nbox_u = nfc > 0 ? nu : nq;
    static constexpr int nterm = 3 + 2 * NQ_, nrow = nbox_u + nx + nfric + NOBS_;
    template <typename T>
    __host__ __device__ static constexpr Layout layout() {
        return compute_layout(LayoutDims{N, nq, nx, nu, neq, nfc, nterm, nrow, nobs, nb, int(sizeof(T))});
    }
};
struct RuntimeDims {
    static constexpr bool kStatic = false;
    static constexpr int nq = 0, nf = 0, nc = 0, nb = 0, nx = 0, nfc = 0, nu = 0, nz = 0, neq = 0, nfric = 0, nbox_u = 0;
    static constexpr int nobs = 0, N = 0, nterm = 0, nrow = 0;
    template <typename T>
    __host__ __device__ static constexpr Layout layout() { return Layout{}; }
};

// TW = warps per instance ("team").  One warp serves the small stage matrices (nz <= 40: everything above is
// written for it); the large ones (cfg3: 84 x 84, cfg5: 68 x 68) take a team of four warps per instance — same
// shared-memory footprint, four times the lanes in every strided loop, named barriers instead of warp barriers.
template <typename T, typename D, int TW = 1>
struct Solver {
    static constexpr int kTS = TW * WARP;   // threads per instance = stride of the lane-parallel loops
    const DevProblem<T>& P;   // shared-memory copy: arrays indexed per lane
    const DevProblem<T>& C;   // kernel-parameter copy (constant bank): scalars and uniformly indexed entries
    const Layout& L;
    const int lane;
#define UB_DIM(FN, name) \
    __device__ __forceinline__ int FN() const { if constexpr (D::kStatic) return D::name; else return P.name; }
    UB_DIM(NQ, nq) UB_DIM(NX, nx) UB_DIM(NU, nu) UB_DIM(NZ, nz) UB_DIM(NFC, nfc) UB_DIM(NEQ, neq) UB_DIM(NFRIC, nfric)
    UB_DIM(NBOXU, nbox_u) UB_DIM(NB, nb) UB_DIM(NC, nc) UB_DIM(NF, nf) UB_DIM(NN, N) UB_DIM(NOBS, nobs)
#undef UB_DIM
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
    __device__ __forceinline__ void place_dynamic_spheres(int k, T ao, T* sph) const {
        for (int s = 0; s < P.nsph; ++s)
            if (P.slink[s] <= -2) {
                const int o = (k * P.ndyn + (-2 - P.slink[s])) * 9;
                for (int c = 0; c < 3; ++c) sph[3 * s + c] = ws[oXO() + o + c] + (ao != T(0) ? ao * ws[oDXO() + o + c] : T(0));
            }
    }
    // the inertial-alignment cost likewise (run-time-dimension kernel only)
    __device__ __forceinline__ bool IALIGN() const { if constexpr (D::kStatic) return false; else return P.iacost != 0; }
    __device__ __forceinline__ int NPAIRS() const { if constexpr (D::kStatic) return D::nobs; else return P.npairs; }
    // workspace / shared-memory offsets: immediates for the specialised kernels
#define UB_OFF(name) \
    __device__ __forceinline__ int o##name() const { if constexpr (D::kStatic) { constexpr Layout l = D::template layout<T>(); return l.name; } else return L.name; }
    UB_OFF(Z) UB_OFF(DZ) UB_OFF(GAP) UB_OFF(LG) UB_OFF(LCT) UB_OFF(LR) UB_OFF(LJP) UB_OFF(LHO) UB_OFF(LJO) UB_OFF(DF)
    UB_OFF(RHOE) UB_OFF(YE) UB_OFF(RHOT) UB_OFF(YT) UB_OFF(TT) UB_OFF(LAM) UB_OFF(FAC) UB_OFF(WF) UB_OFF(XN) UB_OFF(UN)
    UB_OFF(XW) UB_OFF(UW) UB_OFF(TG) UB_OFF(BD) UB_OFF(LIA) UB_OFF(LJA) UB_OFF(XO) UB_OFF(DXO)
#undef UB_OFF
    __device__ __forceinline__ int LDM() const { return NZ() | 1; }
    __device__ __forceinline__ int LDF() const { return NU() | 1; }
    __device__ __forceinline__ int NROW() const { return NBOXU() + NX() + NFRIC() + NOBS(); }
    __device__ __forceinline__ int NTERM() const { return 3 + 2 * NQ(); }
    // 16-byte aligned factor blocks (same rule as compute_layout)
    __device__ __forceinline__ int FSTRIDE() const {
        const int fs_row = (NZ() * LDF() + 3) & ~3, fs_col = (NU() * (NZ() + 1) + 3) & ~3;
        return fs_row > fs_col ? fs_row : fs_col;
    }
    // entry (i, j) of a stored factor block [L; Y]: column-major (column length nz+1) for the blocked kernels
    // — written straight from the panel registers with coalesced stores, read conflict-free by every sweep —
    // row-major otherwise
    static constexpr bool kBlocked = TW == 1 && D::kStatic && D::nu <= 16 && D::nz < 2 * WARP;
    __device__ __forceinline__ int fidx(int i, int j) const {
        if constexpr (kBlocked) return j * (D::nz + 1) + i;
        else return i * LDF() + j;
    }
    // batch data of this instance
    T* ws;   // the one per-instance base address; X, U, target, body are instance-local blocks of it
    T* X;
    T* U;
    const T* target;
    const T* body;
    // shared memory of this warp
    T* sM;
    T* sP;
    T* sPv;
    T* sSA;
    T* sV;
    T* sTT;
    T* sSm;
    T* sRed;     // team reductions / broadcasts (TW > 1)
    int bar_id;  // named barrier of this team (TW > 1)

    // ---- team primitives: plain warp operations for TW = 1 ----
    __device__ __forceinline__ void tsync() const {
        if constexpr (TW == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TW * WARP) : "memory");
    }
    template <typename Op>
    __device__ __forceinline__ T treduce(T v, Op op) const {
        if constexpr (TW > 1) {
            tsync();                                  // the previous reduction has been read by everybody
            if ((lane & (WARP - 1)) == 0) sRed[lane / WARP] = v;
            tsync();
            v = sRed[0];
#pragma unroll
            for (int w = 1; w < TW; ++w) v = op(v, sRed[w]);
        }
        return v;
    }
    __device__ __forceinline__ T tsum(T v) const { return treduce(warp_sum(v), [](T a, T b) { return a + b; }); }
    __device__ __forceinline__ T tmax(T v) const { return treduce(warp_max(v), [](T a, T b) { return max(a, b); }); }
    __device__ __forceinline__ T tmin(T v) const { return treduce(warp_min(v), [](T a, T b) { return min(a, b); }); }
    int nan_reason = 0;  // where a solve first went non-finite: 1 factorisation, 2 step, 4 complementarity, 5 step size, 3 iterate
    // phase cycle counters (profile mode, option stop_after = 9): linearise, factor pass, forward predictor, corrector
    // passes, line search; and inside the factor pass: gradient, matrix build, dynamics terms, factorisation
    long long t_lin = 0, t_fac = 0, t_swp = 0, t_side = 0, t_ls = 0;
    long long t_g = 0, t_f1 = 0, t_f2 = 0, t_f3 = 0;

    __device__ Solver(const DevProblem<T>& P_, const DevProblem<T>& C_, const Layout& L_, int lane_)
        : P(P_), C(C_), L(L_), lane(lane_) {}

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
    __device__ __forceinline__ T row_eps(int fam) const { return row_soft(fam) ? C.invZ : C.eps_hard; }
    // friction row coefficients on the 3 force components of contact c
    __device__ __forceinline__ V3<T> fric_coeff(int c, int which) const {
        const V3<T> n = ld3(P.cn[c]), s0 = ld3(P.cspan[c]), s1 = ld3(P.cspan[c] + 3);
        if (which == 0) return n;
        const T a = (which >= 3) ? T(1) : T(-1), b = (which == 2 || which == 4) ? T(1) : T(-1);
        return P.cmu[c] * n + a * s0 + b * s1;
    }
    // value of ineq row r of stage k at the QP iterate z (stage vector zk = [du; dx]),
    // and bounds.  Uses the linearisation stored in the workspace.
    __device__ T row_value(int k, int r, int fam, const T* zk, const T* xk, const T* uk, T* lb, T* ub) const {
        const int nq = NQ(), nu = NU();
        if (fam == 0) {
            const T u = uk[r];
            *lb = (r < nq ? P.ulb[r] : C.flb) - u;
            *ub = (r < nq ? P.uub[r] : C.fub) - u;
            return zk[r];
        }
        if (fam == 1) {
            const int i = r - NBOXU();
            const T x = xk[i];
            *lb = P.xlb[i] - x;
            *ub = P.xub[i] - x;
            return zk[nu + i];
        }
        *lb = T(0);
        *ub = tinf<T>();
        if (fam == 2) {
            const int i = r - NBOXU() - NX(), c = i / 5;
            const V3<T> a = fric_coeff(c, i % 5);
            const T* f = uk + nq + 3 * c;
            const T* df = zk + nq + 3 * c;
            return a.x * (f[0] + df[0]) + a.y * (f[1] + df[1]) + a.z * (f[2] + df[2]);
        }
        const int i = r - NBOXU() - NX() - NFRIC();
        const int ow = OBSW();
        const T* J = ws + oLJO() + (k * NOBS() + i) * ow;
        T v = ws[oLHO() + k * NOBS() + i];
        for (int j = 0; j < ow; ++j) v += J[j] * zk[nu + j];
        return v;
    }
    // a_r . d  for a stage direction d = [du; dx]
    __device__ T row_dot(int k, int r, int fam, const T* d) const {
        const int nq = NQ(), nu = NU();
        if (fam == 0) return d[r];
        if (fam == 1) return d[nu + r - NBOXU()];
        if (fam == 2) {
            const int i = r - NBOXU() - NX(), c = i / 5;
            const V3<T> a = fric_coeff(c, i % 5);
            const T* df = d + nq + 3 * c;
            return a.x * df[0] + a.y * df[1] + a.z * df[2];
        }
        const int i = r - NBOXU() - NX() - NFRIC();
        const int ow = OBSW();
        const T* J = ws + oLJO() + (k * NOBS() + i) * ow;
        T v = 0;
        for (int j = 0; j < ow; ++j) v += J[j] * d[nu + j];
        return v;
    }
    // Slack/multiplier record of one inequality row: {t_lo, t_hi, lam_lo, lam_hi} and the step
    // {dt_lo, dt_hi, dlam_lo, dlam_hi}, 8 consecutive values (two 16-byte quads) per row so that a warp
    // reads the rows of a stage with fully coalesced vector loads.
    struct alignas(16) Quad {
        T v[4];
    };
    __device__ __forceinline__ Quad* side_tl(int k, int r) const { return reinterpret_cast<Quad*>(ws + oTT()) + (k * NROW() + r) * 2; }
    __device__ __forceinline__ Quad* side_dd(int k, int r) const { return side_tl(k, r) + 1; }
    // Staging of the side records: the records of the stage a pass visits NEXT are copied into shared memory
    // with cp.async while the current stage computes; readers take them from `recs(k)` ([2r] = {t, lam},
    // [2r+1] = {dt, dlam}).  Writers always store to the workspace.
    static constexpr bool kStageTT = TW == 1 && D::kStatic && D::nrow <= UB_STAGE_ROWS_MAX;
    __device__ __forceinline__ const Quad* recs(int k) const {
        if constexpr (kStageTT) return reinterpret_cast<const Quad*>(sTT);
        else return reinterpret_cast<const Quad*>(ws + oTT()) + k * NROW() * 2;
    }
    __device__ __forceinline__ void cp_async16(void* dst, const void* src) const {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    }
    __device__ __forceinline__ void cp_commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
    template <int NYOUNGER>
    __device__ __forceinline__ void cp_wait() const {
        asm volatile("cp.async.wait_group %0;" ::"n"(NYOUNGER) : "memory");
        tsync();
    }
    // issue (no commit) the copy of the records of stage k; k outside [0, N] issues nothing
    __device__ __forceinline__ void tt_issue(int k) const {
        if constexpr (kStageTT) {
            if (k < 0 || k > NN()) return;
            constexpr int CH = NROW_STATIC * 8 * int(sizeof(T)) / 16;
            const char* src = reinterpret_cast<const char*>(ws + oTT() + k * NROW() * 8);
            char* dst = reinterpret_cast<char*>(sTT);
            for (int i = lane; i < CH; i += kTS) cp_async16(dst + 16 * i, src + 16 * i);
        }
    }
    static constexpr int NROW_STATIC = D::nrow;
    // staged per-stage vectors (same schedule as the side records): the QP iterate z_k (or, in the corrector
    // pass, the stored predictor gradient), the linearisation point x_k, u_k, the position Jacobian and w_k
    static constexpr int kSmX = (D::nz + 3) / 4 * 4, kSmU = kSmX + (D::nx + 3) / 4 * 4, kSmJ = kSmU + (D::nu + 3) / 4 * 4,
                         kSmW = kSmJ + (3 * D::nq + 3) / 4 * 4;
    __device__ __forceinline__ const T* st_z(int k) const { if constexpr (kStageTT) return sSm; else return Zk(k); }
    __device__ __forceinline__ const T* st_gp(int k) const { if constexpr (kStageTT) return sSm; else return ws + oLAM() + k * NZ(); }
    __device__ __forceinline__ const T* st_x(int k) const { if constexpr (kStageTT) return sSm + kSmX; else return X + k * NX(); }
    __device__ __forceinline__ const T* st_u(int k) const { if constexpr (kStageTT) return sSm + kSmU; else return U + k * NU(); }
    __device__ __forceinline__ const T* st_jp(int k) const { if constexpr (kStageTT) return sSm + kSmJ; else return ws + oLJP() + k * 3 * NQ(); }
    __device__ __forceinline__ const T* st_w(int k) const { if constexpr (kStageTT) return sSm + kSmW; else return ws + oWF() + k * NU(); }
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
        if constexpr (kStageTT) {
            if (k < 0 || k > NN()) return;
            cp_async_elems(sSm, gp ? ws + oLAM() + k * NZ() : Zk(k), NZ());
            if (!gp) {
                cp_async_elems(sSm + kSmX, X + k * NX(), NX());
                if (k < NN()) cp_async_elems(sSm + kSmU, U + k * NU(), NU());
            }
            if (jp) cp_async_elems(sSm + kSmJ, ws + oLJP() + k * 3 * NQ(), 3 * NQ());
        }
    }
    __device__ __forceinline__ void w_issue(int k) const {
        if constexpr (kStageTT) {
            if (k < 0 || k >= NN()) return;
            cp_async_elems(sSm + kSmW, ws + oWF() + k * NU(), NU());
        }
    }
    // Newton data of one side: returns the barrier weight and the coefficient that multiplies sgn*a in the
    // stage gradient;  d = signed distance to the bound at the current iterate
    __device__ __forceinline__ T side_coef(T t, T lam, T d, T eps, T target, T corr) const {
        const T rd = d + eps * lam - t;
        const T rc = t * lam - target + corr;
        return -lam + fdiv(rc + lam * rd, t + eps * lam);
    }
    // stage vectors are stored with stride nz as [du (nu); dx (nx)]; the terminal
    // stage uses the same slots (its du part is unused and kept at zero)
    __device__ __forceinline__ T* Zk(int k) const { return ws + oZ() + k * NZ(); }
    __device__ __forceinline__ T* DZk(int k) const { return ws + oDZ() + k * NZ(); }

    // number of equality rows of stage k and their data
    __device__ __forceinline__ int neq_of(int k) const { return k < NN() ? NEQ() : NTERM(); }

    // -------------------------------------------------------- linearisation
    // Df: d g / d f, constant in x (compute_object_wrenches, contact_constraints.h:106-157)
    __device__ void build_Df() {
        const T scale = rsqrt(T(6 * NB()));
        T* Df = ws + oDF();
        for (int idx = lane; idx < NEQ() * NFC(); idx += kTS) Df[idx] = T(0);
        tsync();
        for (int j = lane; j < NFC(); j += kTS) {
            const int c = j / NF(), comp = j % NF();
            V3<T> e;
            if (NF() == 1) e = ld3(P.cn[c]);
            else e = V3<T>(comp == 0 ? T(1) : T(0), comp == 1 ? T(1) : T(0), comp == 2 ? T(1) : T(0));
            const int b1 = P.cb1[c], b2 = P.cb2[c];
            if (b1 >= 0) {
                const BodyP<T> Bd = load_body(body + b1 * UB_BODY_PARAMS);
                const V3<T> tq = cross(ld3(P.cr1[c]) - Bd.com, e);
                const T s = -scale / Bd.m;
                Df[(6 * b1 + 0) * NFC() + j] = s * e.x;
                Df[(6 * b1 + 1) * NFC() + j] = s * e.y;
                Df[(6 * b1 + 2) * NFC() + j] = s * e.z;
                Df[(6 * b1 + 3) * NFC() + j] = s * tq.x;
                Df[(6 * b1 + 4) * NFC() + j] = s * tq.y;
                Df[(6 * b1 + 5) * NFC() + j] = s * tq.z;
            }
            {
                const BodyP<T> Bd = load_body(body + b2 * UB_BODY_PARAMS);
                const V3<T> tq = cross(ld3(P.cr2[c]) - Bd.com, T(-1) * e);
                const T s = -scale / Bd.m;
                Df[(6 * b2 + 0) * NFC() + j] = -s * e.x;
                Df[(6 * b2 + 1) * NFC() + j] = -s * e.y;
                Df[(6 * b2 + 2) * NFC() + j] = -s * e.z;
                Df[(6 * b2 + 3) * NFC() + j] = s * tq.x;
                Df[(6 * b2 + 4) * NFC() + j] = s * tq.y;
                Df[(6 * b2 + 5) * NFC() + j] = s * tq.z;
            }
        }
        tsync();
    }

    // Linearise every knot around (X, U): lane j carries d/dx_j.
    // Writes LG [k][neq] (g value incl. Df f), LCT [k][neq][nz] (rows [0 | Df | C] over the stage vector),
    // LR [k][3], LJP [k][3][nq], LHO [k][nobs], LJO [k][nobs][nq], GAP [k][nx].
    __device__ void linearize() {
        const int nq = NQ(), nx = NX(), nu = NU(), N = NN();
        const T scale = rsqrt(T(6 * max(NB(), 1)));
        T sph[3 * UB_MAX_SPHERES], dsph[3 * UB_MAX_SPHERES];
        for (int k = 0; k <= N; ++k) {
            const T* x = X + k * nx;
            Kin<T> Kn;
            KinTan<T> Dt;
            forward_kinematics<T, true>(P, x, lane, Kn, Dt, SPHERES() ? sph : nullptr, dsph);
            if (NXO() > 0) place_dynamic_spheres(k, T(0), sph);
            if (lane == 0) {
                ws[oLR() + 3 * k] = Kn.r.x;
                ws[oLR() + 3 * k + 1] = Kn.r.y;
                ws[oLR() + 3 * k + 2] = Kn.r.z;
            }
            if (lane < nq) {
                T* Jp = ws + oLJP() + k * 3 * nq;
                Jp[lane] = Dt.r.x;
                Jp[nq + lane] = Dt.r.y;
                Jp[2 * nq + lane] = Dt.r.z;
            }
            if (k < N && NEQ() > 0) {
                for (int b = 0; b < NB(); ++b) {
                    const BodyP<T> Bd = load_body(body + b * UB_BODY_PARAMS);
                    T g6[6], dg6[6];
                    object_dynamics_state_part<T, true>(P, Bd, Kn, Dt, scale, g6, dg6);
                    if (lane < nx) {
                        // row-major over the stage vector [du; dx]: lanes write consecutive addresses
                        T* R = ws + oLCT() + (k * NEQ() + 6 * b) * NZ() + nu + lane;
#pragma unroll
                        for (int i = 0; i < 6; ++i) R[i * NZ()] = dg6[i];
                    }
                    for (int idx = lane; idx < 6 * nu; idx += kTS) {
                        const int i = idx / nu, j = idx % nu;
                        ws[oLCT() + (k * NEQ() + 6 * b + i) * NZ() + j] = (j >= nq) ? ws[oDF() + (6 * b + i) * NFC() + (j - nq)] : T(0);
                    }
                    if (lane < 6) {
                        // g = state part + Df f
                        T gv = g6[0];
#pragma unroll
                        for (int i = 1; i < 6; ++i) gv = (lane == i) ? g6[i] : gv;
                        const T* Dfr = ws + oDF() + (6 * b + lane) * NFC();
                        const T* f = U + k * nu + nq;
                        for (int j = 0; j < NFC(); ++j) gv += Dfr[j] * f[j];
                        ws[oLG() + k * NEQ() + 6 * b + lane] = gv;
                    }
                }
            }
            if (IALIGN() && k < N) {
                // e = S C_we' (a - g) / |g| and its Jacobian (inertial_alignment.cpp:151-163)
                T e2[2], de2[2];
                inertial_alignment_error<T, true>(P, Kn, Dt, e2, de2);
                if (lane == 0) {
                    ws[oLIA() + 2 * k] = e2[0];
                    ws[oLIA() + 2 * k + 1] = e2[1];
                }
                if (lane < nx) {
                    ws[oLJA() + (2 * k) * nx + lane] = de2[0];
                    ws[oLJA() + (2 * k + 1) * nx + lane] = de2[1];
                }
            }
            if (NOBS() > 0) {
                for (int i = 0; i < NPAIRS(); ++i) {
                    const int a = P.pa[i], bb = P.pb[i];
                    const V3<T> d(sph[3 * a] - sph[3 * bb], sph[3 * a + 1] - sph[3 * bb + 1], sph[3 * a + 2] - sph[3 * bb + 2]);
                    const T dist = sqrt(dot(d, d));
                    const V3<T> dd(dsph[3 * a] - dsph[3 * bb], dsph[3 * a + 1] - dsph[3 * bb + 1],
                                   dsph[3 * a + 2] - dsph[3 * bb + 2]);
                    T shift = T(0);
                    if (NXO() > 0) {
                        // known Newton step of the obstacle positions: h + (dh/dc_a) dp_a + (dh/dc_b) dp_b, dh/dc = +-d/|d|
                        for (int side = 0; side < 2; ++side) {
                            const int s = side == 0 ? a : bb;
                            if (P.slink[s] > -2) continue;
                            const T* dp = ws + oDXO() + (k * P.ndyn + (-2 - P.slink[s])) * 9;
                            const T proj = (d.x * dp[0] + d.y * dp[1] + d.z * dp[2]) / dist;
                            shift += side == 0 ? proj : -proj;
                        }
                    }
                    if (lane == 0) ws[oLHO() + k * NOBS() + i] = dist - (P.srad[a] + P.srad[bb] + C.dmin) + shift;
                    if (lane < OBSW()) ws[oLJO() + (k * NOBS() + i) * OBSW() + lane] = lane < nq ? dot(d, dd) / dist : T(0);
                }
                if (EEBOX()) {
                    // rows npairs..+2: r_d + upper - r >= 0; rows npairs+3..+5: r - r_d - lower >= 0
                    // (end_effector_box_constraint.h:46-76)
                    const T* tg = target + 3 * k;
                    for (int c = 0; c < 3; ++c) {
                        const int iu = NPAIRS() + c, il = NPAIRS() + 3 + c;
                        if (lane == 0) {
                            ws[oLHO() + k * NOBS() + iu] = tg[c] + P.eb_hi[c] - Kn.r[c];
                            ws[oLHO() + k * NOBS() + il] = Kn.r[c] - tg[c] - P.eb_lo[c];
                        }
                        if (lane < OBSW()) {
                            ws[oLJO() + (k * NOBS() + iu) * OBSW() + lane] = lane < nq ? -Dt.r[c] : T(0);
                            ws[oLJO() + (k * NOBS() + il) * OBSW() + lane] = lane < nq ? Dt.r[c] : T(0);
                        }
                    }
                }
                if (IACON()) {
                    // five inertial-alignment rows (inertial_alignment.cpp:7-53) behind the box rows, dense over x
                    T h5[5], dh5[5];
                    inertial_alignment_rows<T, true>(P, Kn, Dt, h5, dh5);
                    const int i0 = NPAIRS() + (EEBOX() ? 6 : 0);
#pragma unroll
                    for (int r = 0; r < 5; ++r) {
                        if (lane == 0) ws[oLHO() + k * NOBS() + i0 + r] = h5[r];
                        if (lane < nx) ws[oLJO() + (k * NOBS() + i0 + r) * OBSW() + lane] = dh5[r];
                    }
                }
                if (NPROJ() > 0) {
                    // projectile-path rows (projectile_path_constraint.h:108-146) close the family: dense over q with
                    // the time of closest approach held fixed; the obstacle-state block - w s n' [I, t I, t^2/2 I]
                    // meets the known Newton step of the (last) obstacle and lands in the constant
                    const int i0 = NPAIRS() + (EEBOX() ? 6 : 0) + (IACON() ? 5 : 0);
                    const int o = (k * P.ndyn + P.ndyn - 1) * 9;
                    const T* xo = ws + oXO() + o;
                    const T* dxo = ws + oDXO() + o;
                    for (int i = 0; i < NPROJ(); ++i) {
                        const int a = P.proj_sph[i];
                        V3<T> n;
                        T tc;
                        const T h = projectile_row(P, i, ld3(sph + 3 * a), xo, &n, &tc);
                        const T w = C.proj_scale / P.proj_d[i] * C.proj_s;
                        const V3<T> dstep = ld3(dxo) + tc * ld3(dxo + 3) + (T(0.5) * tc * tc) * ld3(dxo + 6);
                        if (lane == 0) ws[oLHO() + k * NOBS() + i0 + i] = h - w * dot(n, dstep);
                        if (lane < OBSW())
                            ws[oLJO() + (k * NOBS() + i0 + i) * OBSW() + lane] = lane < nq ? w * dot(n, ld3(dsph + 3 * a)) : T(0);
                    }
                }
            }
            // dynamics gap b_k = A x_k + B u_k - x_{k+1}  (exact triple integrator, system_dynamics.h:15-26)
            if (k < N && lane < nq) {
                const T dt = C.dt;
                const T* xn = X + (k + 1) * nx;
                const T q = x[lane], v = x[nq + lane], a = x[2 * nq + lane], j = U[k * nu + lane];
                T* gap = ws + oGAP() + k * nx;
                gap[lane] = q + dt * v + T(0.5) * dt * dt * a + dt * dt * dt / T(6) * j - xn[lane];
                gap[nq + lane] = v + dt * a + T(0.5) * dt * dt * j - xn[nq + lane];
                gap[2 * nq + lane] = a + dt * j - xn[2 * nq + lane];
            }
        }
        tsync();
    }

    // --------------------------------------------------- performance index
    // Lane k evaluates knot k (values only).  Mirrors orc::performance().
    // `ao`: step along the (known) obstacle-state direction, 0 for the current iterate
    __device__ Perf<T> performance(const T* Xt, const T* Ut, T ao = T(0)) const {
        const int nq = NQ(), nx = NX(), nu = NU(), N = NN();
        const T dt = C.dt;
        const T scale = rsqrt(T(6 * max(NB(), 1)));
        T cost = 0, dyn = 0, eq = 0, ineq = 0, max_eq = 0, min_margin = tinf<T>();
        T sph[3 * UB_MAX_SPHERES];
        for (int k = lane; k <= N; k += kTS) {
            const T* x = Xt + k * nx;
            Kin<T> Kn;
            KinTan<T> Dn;
            forward_kinematics<T, false>(P, x, -1, Kn, Dn, SPHERES() ? sph : nullptr, nullptr);
            if (NXO() > 0) {
                place_dynamic_spheres(k, ao, sph);
                if (k < N)   // dynamics defect of the obstacle states: (1 - ao) x the defect of the iterate
                    for (int j = 0; j < P.ndyn; ++j) {
                        const T* o0 = ws + oXO() + (k * P.ndyn + j) * 9;
                        const T* o1 = ws + oXO() + ((k + 1) * P.ndyn + j) * 9;
                        for (int c = 0; c < 3; ++c) {
                            const T g0 = o0[c] + dt * o0[3 + c] + T(0.5) * dt * dt * o0[6 + c] - o1[c];
                            const T g1 = o0[3 + c] + dt * o0[6 + c] - o1[3 + c];
                            const T g2 = o0[6 + c] - o1[6 + c];
                            dyn += dt * (T(1) - ao) * (T(1) - ao) * (g0 * g0 + g1 * g1 + g2 * g2);
                        }
                    }
            }
            const T* rd = target + 3 * k;
            if (k == N) {
                for (int i = 0; i < 3; ++i) {
                    const T e = rd[i] - Kn.r[i];
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
                    const T lo = x[i] - P.xlb[i], hi = P.xub[i] - x[i];
                    const T a = min(T(0), lo), b = min(T(0), hi);
                    ineq += dt * (a * a + b * b);
                    min_margin = min(min_margin, min(lo, hi));
                }
            if (k == N) continue;
            const T* u = Ut + k * nu;
            T c = 0;
            for (int i = 0; i < nx; ++i) {
                const T e = x[i] - P.xd[i];
                c += T(0.5) * P.Qd[i] * e * e;
            }
            for (int i = 0; i < nq; ++i) c += T(0.5) * P.Rd[i] * u[i] * u[i];
            for (int i = 0; i < NFC(); ++i) c += T(0.5) * C.fw * u[nq + i] * u[nq + i];
            for (int i = 0; i < 3; ++i) {
                const T e = Kn.r[i] - rd[i];
                c += T(0.5) * P.Wd[i] * e * e;
            }
            if (IALIGN()) {
                T e2[2];
                inertial_alignment_error<T, false>(P, Kn, Dn, e2, nullptr);
                c += T(0.5) * P.ia_w * (e2[0] * e2[0] + e2[1] * e2[1]);
            }
            cost += dt * c;
            const T* xn = Xt + (k + 1) * nx;
            for (int i = 0; i < nq; ++i) {
                const T q = x[i], v = x[nq + i], a = x[2 * nq + i], j = u[i];
                const T g0 = q + dt * v + T(0.5) * dt * dt * a + dt * dt * dt / T(6) * j - xn[i];
                const T g1 = v + dt * a + T(0.5) * dt * dt * j - xn[nq + i];
                const T g2 = a + dt * j - xn[2 * nq + i];
                dyn += dt * (g0 * g0 + g1 * g1 + g2 * g2);
            }
            const int nbox = NBOXU();
            for (int i = 0; i < nbox; ++i) {
                const T lo = u[i] - (i < nq ? P.ulb[i] : C.flb), hi = (i < nq ? P.uub[i] : C.fub) - u[i];
                const T a = min(T(0), lo), b = min(T(0), hi);
                ineq += dt * (a * a + b * b);
                min_margin = min(min_margin, min(lo, hi));
            }
            for (int b = 0; b < (NEQ() > 0 ? NB() : 0); ++b) {
                const BodyP<T> Bd = load_body(body + b * UB_BODY_PARAMS);
                T g6[6];
                object_dynamics_state_part<T, false>(P, Bd, Kn, Dn, scale, g6, nullptr);
                for (int i = 0; i < 6; ++i) {
                    const T* Dfr = ws + oDF() + (6 * b + i) * NFC();
                    T gv = g6[i];
                    for (int j = 0; j < NFC(); ++j) gv += Dfr[j] * u[nq + j];
                    eq += dt * gv * gv;
                    max_eq = max(max_eq, fabs(gv));
                }
            }
            for (int i = 0; i < NFRIC(); ++i) {
                const int cidx = i / 5;
                const V3<T> a = fric_coeff(cidx, i % 5);
                const T* f = u + nq + 3 * cidx;
                const T h = a.x * f[0] + a.y * f[1] + a.z * f[2];
                const T m = min(T(0), h);
                ineq += dt * m * m;
                min_margin = min(min_margin, h);
            }
            if (k >= 1)
            {
                for (int i = 0; i < NPAIRS(); ++i) {
                    const int a = P.pa[i], bb = P.pb[i];
                    const V3<T> d(sph[3 * a] - sph[3 * bb], sph[3 * a + 1] - sph[3 * bb + 1], sph[3 * a + 2] - sph[3 * bb + 2]);
                    const T h = sqrt(dot(d, d)) - (P.srad[a] + P.srad[bb] + C.dmin);
                    const T m = min(T(0), h);
                    ineq += dt * m * m;
                    min_margin = min(min_margin, h);
                }
                if (EEBOX())
                    for (int c = 0; c < 3; ++c) {
                        const T hu = rd[c] + P.eb_hi[c] - Kn.r[c], hl = Kn.r[c] - rd[c] - P.eb_lo[c];
                        const T mu_ = min(T(0), hu), ml = min(T(0), hl);
                        ineq += dt * (mu_ * mu_ + ml * ml);
                        min_margin = min(min_margin, min(hu, hl));
                    }
                if (IACON()) {
                    T h5[5];
                    inertial_alignment_rows<T, false>(P, Kn, Dn, h5, nullptr);
                    for (int r = 0; r < 5; ++r) {
                        const T m = min(T(0), h5[r]);
                        ineq += dt * m * m;
                        min_margin = min(min_margin, h5[r]);
                    }
                }
                if (NPROJ() > 0) {
                    const int o = (k * P.ndyn + P.ndyn - 1) * 9;
                    T xo[9];
                    for (int c = 0; c < 9; ++c) xo[c] = ws[oXO() + o + c] + (ao != T(0) ? ao * ws[oDXO() + o + c] : T(0));
                    for (int i = 0; i < NPROJ(); ++i) {
                        V3<T> n;
                        T tc;
                        const T h = projectile_row(P, i, ld3(sph + 3 * P.proj_sph[i]), xo, &n, &tc);
                        const T m = min(T(0), h);
                        ineq += dt * m * m;
                        min_margin = min(min_margin, h);
                    }
                }
            }
        }
        Perf<T> pf;
        pf.cost = tsum(cost);
        pf.dyn = tsum(dyn);
        pf.eq = tsum(eq);
        pf.ineq = tsum(ineq);
        pf.max_eq = tmax(max_eq);
        pf.min_margin = tmin(min_margin);
        return pf;
    }

    // ------------------------------------------------------------ QP pieces
    // Load the equality rows of stage k into shared memory SA [row][nz] in
    // stage-vector index order, with their constant c, penalty rho and
    // multiplier y in sV-side arrays (global RHOE/YE, RHOT/YT).
    __device__ void load_eq_rows(int k) {
        const int nq = NQ(), nx = NX(), nu = NU(), nz = NZ();
        if (k < NN()) {
            const T* __restrict__ R = ws + oLCT() + k * NEQ() * nz;
            for (int idx = lane; idx < NEQ() * nz; idx += kTS) sSA[idx] = R[idx];
        } else {
            // terminal equality [r_d - r; v; a] = 0: three dense rows over q, the rest are unit rows
            for (int idx = lane; idx < 3 * nz; idx += kTS) {
                const int i = idx / nz, j = idx % nz;
                T v = T(0);
                if (j >= nu && j < nu + nq) v = -ws[oLJP() + (k * 3 + i) * nq + (j - nu)];
                sSA[idx] = v;
            }
        }
        tsync();
    }
    // constant (value at z = 0) of equality row i of stage k
    __device__ __forceinline__ T eq_const(int k, int i) const {
        if (k < NN()) return ws[oLG() + k * NEQ() + i];
        if (i < 3) return target[3 * k + i] - ws[oLR() + 3 * k + i];
        return X[k * NX() + NQ() + (i - 3)];
    }
    __device__ __forceinline__ T* rho_eq(int k) const { return k < NN() ? ws + oRHOE() + k * NEQ() : ws + oRHOT(); }
    __device__ __forceinline__ T* y_eq(int k) const { return k < NN() ? ws + oYE() + k * NEQ() : ws + oYT(); }
    // value a_i . z + c of equality row i (dense rows from SA; terminal unit rows direct)
    __device__ T eq_value(int k, int i, const T* zk) const {
        if (k == NN() && i >= 3) return zk[NU() + NQ() + (i - 3)] + eq_const(k, i);
        const T* a = sSA + i * NZ();
        T v = eq_const(k, i);
        for (int j = (k < NN() ? NQ() : NU()); j < NZ(); ++j) v += a[j] * zk[j];
        return v;
    }

    // Set the proximal weights of the equality rows (soft: Z; hard: rho_hard on the
    // unit-normalised row) and reset the multipliers.
    __device__ void init_eq_weights() {
        for (int k = 0; k <= NN(); ++k) {
            const int ne = neq_of(k);
            if (ne == 0) continue;
            load_eq_rows(k);
            for (int i = lane; i < ne; i += kTS) {
                T rho = C.Z;
                if (!C.soft_poly) {
                    T n2 = T(1);
                    if (!(k == NN() && i >= 3)) {
                        n2 = T(0);
                        for (int j = 0; j < NZ(); ++j) n2 += sSA[i * NZ() + j] * sSA[i * NZ() + j];
                    }
                    rho = n2 > T(0) ? C.rho_hard / n2 : T(0);
                }
                rho_eq(k)[i] = rho;
                y_eq(k)[i] = T(0);
            }
            tsync();
        }
    }

    // M (lower triangle, ld = LDM()) += [B A]' Pn [B A] with the block structure
    // A = A3 (x) I, B = B3 (x) I of the exact triple-integrator discretisation.
    // ASSIGN: the stage matrix has not been initialised — every (I >= J) block entry is assigned and the force
    // rows / columns, which the dynamics do not touch, are zeroed (saves the separate zero fill of the buffer)
    template <bool ASSIGN = false>
    __device__ void add_dynamics_hessian() {
        const int nq = NQ(), nu = NU(), nx = NX(), ld = LDM();
        const T dt = C.dt;
        // T3[a][I]: column 0 = B3, columns 1..3 = A3 (block order of the stage vector: jerk, q, v, a)
        const T T3[3][4] = {{dt * dt * dt / T(6), T(1), dt, T(0.5) * dt * dt},
                            {T(0.5) * dt * dt, T(0), T(1), dt},
                            {dt, T(0), T(0), T(1)}};
        if constexpr (D::kStatic) {
            // One pass over the nq x nq entry positions (rolled: the body stays in the instruction cache): the nine
            // P_ab(ii, jj) are loaded once and feed all ten block pairs (I >= J); zero entries of T3 drop out at
            // compile time (52 products per position).
#pragma unroll 1
            for (int e = lane; e < D::nq * D::nq; e += kTS) {
                const int ii = e / D::nq, jj = e % D::nq;
                T pab[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) pab[a][b] = sP[(a * D::nq + ii) * D::nx + b * D::nq + jj];
#pragma unroll
                for (int I = 0; I < 4; ++I) {
#pragma unroll
                    for (int J = 0; J <= I; ++J) {
                        T acc = T(0);
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            if ((I == 1 && a != 0) || (I == 2 && a == 2)) continue;
#pragma unroll
                            for (int b = 0; b < 3; ++b) {
                                if ((J == 1 && b != 0) || (J == 2 && b == 2)) continue;
                                acc += T3[a][I] * T3[b][J] * pab[a][b];
                            }
                        }
                        const int mi = (I == 0) ? ii : D::nu + (I - 1) * D::nq + ii;
                        const int mj = (J == 0) ? jj : D::nu + (J - 1) * D::nq + jj;
                        if constexpr (ASSIGN) sM[mi * ld + mj] = acc;
                        else sM[mi * ld + mj] += acc;  // diagonal blocks also touch (unused) upper entries
                    }
                }
            }
            if constexpr (ASSIGN && D::nfc > 0) {
                // force rows (all columns up to the diagonal) and force columns of the state rows
                for (int idx = lane; idx < D::nfc * D::nu; idx += kTS) sM[(D::nq + idx / D::nu) * ld + idx % D::nu] = T(0);
                for (int idx = lane; idx < D::nx * D::nfc; idx += kTS) sM[(D::nu + idx / D::nfc) * ld + D::nq + idx % D::nfc] = T(0);
            }
        } else {
            const int nb4 = 4 * nq;
            for (int idx = lane; idx < nb4 * nb4; idx += kTS) {
                const int bi = idx / nb4, bj = idx % nb4;
                if (bj > bi) continue;
                const int I = bi / nq, ii = bi % nq, J = bj / nq, jj = bj % nq;
                T acc = T(0);
                for (int a = 0; a < 3; ++a) {
                    const T ta = T3[a][I];
                    if (ta == T(0)) continue;
                    for (int b = 0; b < 3; ++b) {
                        const T tb = T3[b][J];
                        if (tb == T(0)) continue;
                        acc += ta * tb * sP[(a * nq + ii) * nx + b * nq + jj];
                    }
                }
                const int mi = (I == 0) ? ii : nu + (I - 1) * nq + ii;
                const int mj = (J == 0) ? jj : nu + (J - 1) * nq + jj;
                sM[mi * ld + mj] += acc;
            }
        }
        tsync();
    }
    // vec (stage layout) += [B A]' pv
    __device__ void add_dynamics_gradient(T* vec) const {
        const int nq = NQ(), nu = NU();
        const T dt = C.dt;
        if (lane < nq) {
            const T p0 = sPv[lane], p1 = sPv[nq + lane], p2 = sPv[2 * nq + lane];
            vec[lane] += dt * dt * dt / T(6) * p0 + T(0.5) * dt * dt * p1 + dt * p2;
            vec[nu + lane] += p0;
            vec[nu + nq + lane] += dt * p0 + p1;
            vec[nu + 2 * nq + lane] += T(0.5) * dt * dt * p0 + dt * p1 + p2;
        }
        tsync();
    }

    // Build the Newton matrix of stage k in sM (lower triangle): cost Hessian +
    // equality proximal terms + barrier terms of the inequality sides.
    // `initialised`: the dynamics term has already been ASSIGNED to the buffer (add_dynamics_hessian<true>)
    __device__ void build_stage_matrix(int k, bool rows_loaded = false, bool initialised = false) {
        const int nq = NQ(), nu = NU(), nx = NX(), nz = NZ(), ld = LDM();
        const T dt = C.dt;
        if (!initialised) {
            for (int idx = lane; idx < nz * ld; idx += kTS) sM[idx] = T(0);
            tsync();
        }
        if (k < NN()) {
            // cost (quadratic_joint_state_input_cost.h:9-33, end_effector_cost.h:48-84), scaled by dt
            for (int i = lane; i < nz; i += kTS) {
                T d;
                if (i < nq) d = dt * P.Rd[i] + C.reg_input;
                else if (i < nu) d = dt * C.fw + C.reg_input;
                else d = dt * P.Qd[i - nu];
                sM[i * ld + i] += d;
            }
            tsync();
            const T* Jp = st_jp(k);
            for (int idx = lane; idx < nq * nq; idx += kTS) {
                const int a = idx / nq, b = idx % nq;
                if (b > a) continue;
                T acc = 0;
                for (int c = 0; c < 3; ++c) acc += C.Wd[c] * Jp[c * nq + a] * Jp[c * nq + b];
                sM[(nu + a) * ld + nu + b] += dt * acc;
            }
            tsync();
            if (IALIGN()) {   // Gauss-Newton Hessian w Je' Je of the inertial-alignment cost (dense over x)
                const T* Ja = ws + oLJA() + 2 * k * nx;
                const T wa = dt * P.ia_w;
                for (int idx = lane; idx < nx * nx; idx += kTS) {
                    const int a = idx / nx, b = idx % nx;
                    if (b > a) continue;
                    sM[(nu + a) * ld + nu + b] += wa * (Ja[a] * Ja[b] + Ja[nx + a] * Ja[nx + b]);
                }
                tsync();
            }
        }
        // equality rows: rho a a'
        const int ne = neq_of(k);
        if (ne > 0) {
            if (!rows_loaded) load_eq_rows(k);
            const T* rho = rho_eq(k);
            const int nd = (k < NN()) ? ne : 3;
            const int j0 = (k < NN()) ? nq : nu;  // first column with non-zeros
            const int span = nz - j0;
            if (D::kStatic && D::neq <= 8 && k < NN()) {
                constexpr int NE = D::kStatic ? (D::neq > 0 ? D::neq : 1) : 1;
                for (int c = j0 + lane; c < nz; c += kTS) {
                    T ac[NE];
#pragma unroll
                    for (int r = 0; r < NE; ++r) ac[r] = rho[r] * sSA[r * nz + c];
                    for (int i = c; i < nz; ++i) {
                        T acc = 0;
#pragma unroll
                        for (int r = 0; r < NE; ++r) acc += ac[r] * sSA[r * nz + i];
                        sM[i * ld + c] += acc;
                    }
                }
            } else {
                for (int idx = lane; idx < span * span; idx += kTS) {
                    const int i = j0 + idx / span, j = j0 + idx % span;
                    if (j > i) continue;
                    T acc = 0;
                    for (int r = 0; r < nd; ++r) acc += rho[r] * sSA[r * nz + i] * sSA[r * nz + j];
                    sM[i * ld + j] += acc;
                }
            }
            if (k == NN()) {
                tsync();   // the dense-row loop above adds (zeros) to the same diagonal entries
                for (int i = 3 + lane; i < ne; i += kTS) {
                    const int m = nu + nq + (i - 3);
                    sM[m * ld + m] += rho[i];
                }
            }
            tsync();
        }
        // inequality sides: w a a', w = lam / (t + eps lam)
        const int nbx = NBOXU() + NX();
        for (int r = lane; r < nbx; r += kTS) {
            const int fam = r < NBOXU() ? 0 : 1;
            if (!row_valid(k, fam)) continue;
            const T eps = row_eps(fam);
            const Quad q = recs(k)[2 * r];
            const T w = fdiv(q.v[2], q.v[0] + eps * q.v[2]) + fdiv(q.v[3], q.v[1] + eps * q.v[3]);
            const int m = fam == 0 ? r : nu + (r - NBOXU());
            sM[m * ld + m] += w;
        }
        tsync();
        if (NFRIC() > 0 && k < NN()) {
            const T eps = row_eps(2);
            // one lane per contact: its five pyramid rows give a symmetric 3x3 block
            for (int c = lane; c < NC(); c += kTS) {
                T blk[6] = {0, 0, 0, 0, 0, 0};
                for (int which = 0; which < 5; ++which) {
                    const Quad q = recs(k)[2 * (nbx + 5 * c + which)];
                    const T w = fdiv(q.v[2], q.v[0] + eps * q.v[2]);
                    const V3<T> cf = fric_coeff(c, which);
                    blk[0] += w * cf.x * cf.x;
                    blk[1] += w * cf.y * cf.x;
                    blk[2] += w * cf.y * cf.y;
                    blk[3] += w * cf.z * cf.x;
                    blk[4] += w * cf.z * cf.y;
                    blk[5] += w * cf.z * cf.z;
                }
                T* Mb = sM + (nq + 3 * c) * ld + nq + 3 * c;
                Mb[0] += blk[0];
                Mb[ld] += blk[1];
                Mb[ld + 1] += blk[2];
                Mb[2 * ld] += blk[3];
                Mb[2 * ld + 1] += blk[4];
                Mb[2 * ld + 2] += blk[5];
            }
            tsync();
        }
        if (NOBS() > 0 && k >= 1 && k < NN()) {
            const T eps = row_eps(3);
            T* wrow = sV + 4 * nz;  // barrier weights of the obstacle rows
            for (int i = lane; i < NOBS(); i += kTS) {
                const Quad q = recs(k)[2 * (nbx + NFRIC() + i)];
                wrow[i] = fdiv(q.v[2], q.v[0] + eps * q.v[2]);
            }
            tsync();
            const int ow = OBSW();
            for (int idx = lane; idx < ow * ow; idx += kTS) {
                const int a = idx / ow, b = idx % ow;
                if (b > a) continue;
                T acc = 0;
                for (int i = 0; i < NOBS(); ++i) {
                    const T* J = ws + oLJO() + (k * NOBS() + i) * ow;
                    acc += wrow[i] * J[a] * J[b];
                }
                sM[(nu + a) * ld + nu + b] += acc;
            }
            tsync();
        }
    }

    // Right-looking Cholesky of the first nu columns of sM (n x n, lower, ld):
    // afterwards columns j < nu hold L (diagonal stored INVERTED) and the
    // trailing block holds the Schur complement.  Lane l owns columns l, l+32, ...
    __device__ bool partial_cholesky(int n, int npiv) {
        const int ld = LDM();
        bool ok = true;
        for (int j = 0; j < npiv; ++j) {
            T d = sM[j * ld + j];
            if (d != d) ok = false;
            if (!(d > C.reg_input)) d = C.reg_input;   // see stage_factor_blocked
            const T inv = rsqrt(d);
            for (int i = j + 1 + lane; i < n; i += kTS) sM[i * ld + j] *= inv;
            tsync();
            if (lane == 0) sM[j * ld + j] = inv;
            if constexpr (TW > 1) {
                // team: 2-D cyclic decomposition of the trailing lower triangle (rows over lane / 16, columns over
                // lane % 16) — balanced, where one column per lane would leave most of the 128 lanes idle
                constexpr int CB = 16, RA = kTS / CB;
                const int cb = lane % CB, ra = lane / CB;
                for (int i = j + 1 + ra; i < n; i += RA) {
                    const T mij = sM[i * ld + j];
                    for (int l = j + 1 + cb; l <= i; l += CB) sM[i * ld + l] -= mij * sM[l * ld + j];
                }
            } else
            for (int l = j + 1 + lane; l < n; l += kTS) {
                const T mlj = sM[l * ld + j];
                T* __restrict__ dst = sM + l * ld + l;         // column l, rows l..n-1
                const T* __restrict__ src = sM + l * ld + j;   // column j, rows l..n-1
                int i = l;
                for (; i + 4 <= n; i += 4) {
                    const T s0 = src[0], s1 = src[ld], s2 = src[2 * ld], s3 = src[3 * ld];
                    const T d0 = dst[0], d1 = dst[ld], d2 = dst[2 * ld], d3 = dst[3 * ld];
                    dst[0] = d0 - s0 * mlj;
                    dst[ld] = d1 - s1 * mlj;
                    dst[2 * ld] = d2 - s2 * mlj;
                    dst[3 * ld] = d3 - s3 * mlj;
                    src += 4 * ld;
                    dst += 4 * ld;
                }
                for (; i < n; ++i) {
                    dst[0] -= src[0] * mlj;
                    src += ld;
                    dst += ld;
                }
            }
            tsync();
        }
        return ok;
    }

    // Blocked factorisation of the AUGMENTED stage matrix [M m; m' .] for small input blocks (nu <= 16, nz < 64):
    //   panel   [L; Y; w'] = [M; m'][:, 0:nu] L^{-T}   right-looking, lane = matrix row (rows in registers, the pivot
    //                                                  column travels by warp shuffle), diagonal stored INVERTED;
    //                                                  the gradient rides along as row nz, so the forward
    //                                                  substitution w = L^{-1} m_u costs nothing extra;
    //   Schur   [P p] = [Mxx m_x] - Y [Y; w']'         8 x 4 lane grid, TR x TC accumulator tile per lane; the spare
    //                                                  tile column carries p = m_x - Y w.  P goes straight to sP as
    //                                                  the full symmetric cost-to-go Hessian, p to sPv.
    // The factor block is stored to the workspace column-major from the panel registers (coalesced).
    // On entry vec = stage gradient [m_u; m_x]; on exit vec[0, nu) = w.
    template <int NU_, int NX_>
    __device__ bool stage_factor_blocked(T* vec, T* Fg) {
        constexpr int NZ_ = NU_ + NX_, AUG = NZ_;
        constexpr int R = (NZ_ + 1 + kTS - 1) / kTS;
        constexpr int TR = (NX_ + 7) / 8, TC = (NX_ + 1 + 3) / 4;
        static_assert(NU_ <= 16 && R <= 2 && 4 * TC > NX_, "blocked factorisation is for small stage matrices");
        const int ld = LDM();
        T row[R][NU_];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = lane + kTS * r;
            const T* src = (i == AUG) ? vec : sM + min(i, NZ_ - 1) * ld;
#pragma unroll
            for (int c = 0; c < NU_; ++c) row[r][c] = src[c];
        }
        bool ok = true;
#pragma unroll
        for (int j = 0; j < NU_; ++j) {
            // The exact pivot is bounded below by reg_input (M_uu >= reg_input I); one that roundoff pushed under
            // the bound is raised to it, which keeps the fp32 factorisation finite (the Newton step is then inexact
            // and the interior-point iteration corrects it).  A NaN pivot still fails the solve.
            T d = __shfl_sync(FULL, row[0][j], j);
            if (d != d) ok = false;
            if (!(d > C.reg_input)) d = C.reg_input;
            const T inv = rsqrt(d);
            T lij[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = lane + kTS * r;
                lij[r] = (i > j) ? row[r][j] * inv : T(0);
            }
#pragma unroll
            for (int c = j + 1; c < NU_; ++c) {
                const T lcj = __shfl_sync(FULL, lij[0], c);
#pragma unroll
                for (int r = 0; r < R; ++r) row[r][c] -= lij[r] * lcj;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = lane + kTS * r;
                row[r][j] = (i > j) ? lij[r] : (i == j ? inv : T(0));
            }
        }
        tsync();   // lanes beyond the last row read a clamped (real) row above
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = lane + kTS * r;
            if (i < NZ_) {
#pragma unroll
                for (int c = 0; c < NU_; ++c) {
                    if (i >= NU_) sM[i * ld + c] = row[r][c];   // Y feeds the Schur update
                    Fg[c * (NZ_ + 1) + i] = row[r][c];          // factor block -> workspace (column-major)
                }
            } else if (i == AUG) {
#pragma unroll
                for (int c = 0; c < NU_; ++c) vec[c] = row[r][c];  // w = L^{-1} m_u
            }
        }
        tsync();
        // Schur complement tile of lane (a, b): rows a*TR.., columns b*TC.. of the state block; column NX_ = p
        const int a = lane >> 2, b = lane & 3;
        int ri[TR];
        const T* ycp[TC];
#pragma unroll
        for (int ii = 0; ii < TR; ++ii) ri[ii] = min(a * TR + ii, NX_ - 1);
        T acc[TR][TC];
#pragma unroll
        for (int cc = 0; cc < TC; ++cc) {
            const int c = b * TC + cc;
            const int cl = min(c, NX_ - 1);
            ycp[cc] = (c == NX_) ? vec : sM + (NU_ + cl) * ld;
#pragma unroll
            for (int ii = 0; ii < TR; ++ii) {
                const int hi = max(ri[ii], cl), lo = min(ri[ii], cl);
                acc[ii][cc] = (c == NX_) ? vec[NU_ + ri[ii]] : sM[(NU_ + hi) * ld + NU_ + lo];
            }
        }
#pragma unroll 1
        for (int m = 0; m < NU_; ++m) {   // rolled: 39 instructions that stay in the instruction cache
            T yr[TR], yc[TC];
#pragma unroll
            for (int ii = 0; ii < TR; ++ii) yr[ii] = sM[(NU_ + ri[ii]) * ld + m];
#pragma unroll
            for (int cc = 0; cc < TC; ++cc) yc[cc] = ycp[cc][m];
#pragma unroll
            for (int ii = 0; ii < TR; ++ii)
#pragma unroll
                for (int cc = 0; cc < TC; ++cc) acc[ii][cc] -= yr[ii] * yc[cc];
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

    // generic path: factor the stage matrix in sM; leaves [L; Y] in its first nu columns and the new cost-to-go
    // Hessian in sP
    __device__ __forceinline__ bool stage_cholesky() {
        const bool ok = partial_cholesky(NZ(), NU());
        copy_cost_to_go();
        return ok;
    }
    // cost-to-go Hessian = trailing block of sM, expanded to the full symmetric matrix
    __device__ __forceinline__ void copy_cost_to_go() {
        const int nu = NU(), nx = NX(), ld = LDM();
        for (int idx = lane; idx < nx * nx; idx += kTS) {
            const int i = idx / nx, j = idx % nx;
            sP[idx] = (j <= i) ? sM[(nu + i) * ld + nu + j] : sM[(nu + j) * ld + nu + i];
        }
        tsync();
    }

    // Stage gradient of the barrier/proximal Lagrangian at the current iterate:
    //   H z + g  +  sum_eq a (rho e + y)  +  sum_sides sgn a [ -lam + (rc + lam rd)/(t + eps lam) ]
    // with rc = t lam - target (+ dt_aff dlam_aff in the corrector).  Result in vec (shared).
    // Every row family accumulates into distinct entries per lane (no atomics).
    __device__ void stage_gradient(int k, bool corrector, T mu_target, T* vec, bool rows_loaded = false) {
        const int nq = NQ(), nu = NU(), nx = NX(), nz = NZ();
        const T dt = C.dt;
        const T* zk = st_z(k);
        const T cm = corrector ? T(1) : T(0);
        // cost part (zero at the terminal stage)
        if (k < NN()) {
            const T* x = st_x(k);
            const T* u = st_u(k);
            const T* Jp = st_jp(k);
            // e = Jp dq + r - r_d, reduced over the warp
            T e3[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const T part = lane < nq ? Jp[c * nq + lane] * zk[nu + lane] : T(0);
                e3[c] = tsum(part) + ws[oLR() + 3 * k + c] - target[3 * k + c];
            }
            // inertial-alignment residual at the QP iterate: e + Je dx
            T ea[2] = {T(0), T(0)};
            const T* Ja = ws + oLJA() + 2 * k * nx;
            if (IALIGN()) {
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    T part = T(0);
                    for (int j = lane; j < nx; j += kTS) part += Ja[r * nx + j] * zk[nu + j];
                    ea[r] = tsum(part) + ws[oLIA() + 2 * k + r];
                }
            }
            for (int i = lane; i < nz; i += kTS) {
                T g;
                if (i < nq) g = dt * P.Rd[i] * (u[i] + zk[i]) + C.reg_input * zk[i];
                else if (i < nu) g = dt * C.fw * (u[i] + zk[i]) + C.reg_input * zk[i];
                else {
                    const int xi = i - nu;
                    g = dt * P.Qd[xi] * (x[xi] + zk[i] - P.xd[xi]);
                    if (xi < nq) g += dt * (C.Wd[0] * Jp[xi] * e3[0] + C.Wd[1] * Jp[nq + xi] * e3[1] + C.Wd[2] * Jp[2 * nq + xi] * e3[2]);
                    if (IALIGN()) g += dt * P.ia_w * (Ja[xi] * ea[0] + Ja[nx + xi] * ea[1]);
                }
                vec[i] = g;
            }
        } else {
            for (int i = lane; i < nz; i += kTS) vec[i] = T(0);
        }
        tsync();
        // box rows: one entry each
        for (int r = lane; r < NBOXU() + nx; r += kTS) {
            const int fam = r < NBOXU() ? 0 : 1;
            if (!row_valid(k, fam)) continue;
            const int m = fam == 0 ? r : nu + (r - NBOXU());
            T lb, ub;
            const T val = zk[m];
            if (fam == 0) {
                const T uu = st_u(k)[r];
                lb = (r < nq ? P.ulb[r] : C.flb) - uu;
                ub = (r < nq ? P.uub[r] : C.fub) - uu;
            } else {
                const int i = r - NBOXU();
                const T xx = st_x(k)[i];
                lb = P.xlb[i] - xx;
                ub = P.xub[i] - xx;
            }
            const T eps = row_eps(fam);
            const Quad q = recs(k)[2 * r];
            Quad dd;
            if (corrector) dd = recs(k)[2 * r + 1];
            else dd.v[0] = dd.v[1] = dd.v[2] = dd.v[3] = T(0);
            const T c0 = side_coef(q.v[0], q.v[2], val - lb, eps, mu_target, cm * dd.v[0] * dd.v[2]);
            const T c1 = side_coef(q.v[1], q.v[3], ub - val, eps, mu_target, cm * dd.v[1] * dd.v[3]);
            vec[m] += c0 - c1;
        }
        tsync();
        // equality rows
        const int ne = neq_of(k);
        if (ne > 0) {
            if (!rows_loaded) load_eq_rows(k);
            const T* rho = rho_eq(k);
            const T* y = y_eq(k);
            const int nd = (k < NN()) ? ne : 3;
            T* mrow = sV + 4 * nz;  // per-row multiplier estimate m_i = rho e_i + y_i
            if (D::kStatic && D::neq <= 8 && k < NN()) {
                // few dense rows: every row value as a warp-wide dot product
                for (int i = 0; i < nd; ++i) {
                    T part = T(0);
                    for (int j = nq + lane; j < nz; j += kTS) part += sSA[i * nz + j] * zk[j];
                    const T e = tsum(part) + eq_const(k, i);
                    if (lane == 0) mrow[i] = rho[i] * e + y[i];
                }
            } else {
                for (int i = lane; i < ne; i += kTS) {
                    const T m = rho[i] * eq_value(k, i, zk) + y[i];
                    if (i < nd) mrow[i] = m;
                    else vec[nu + nq + (i - 3)] += m;  // terminal unit rows (distinct entries)
                }
            }
            tsync();
            for (int j = lane; j < nz; j += kTS) {
                T acc = 0;
                for (int i = 0; i < nd; ++i) acc += mrow[i] * sSA[i * nz + j];
                vec[j] += acc;
            }
            tsync();
        }
        const int nbx = NBOXU() + nx;
        if (NFRIC() > 0 && k < NN()) {
            // one lane per contact: five pyramid rows -> three force entries
            const T eps = row_eps(2);
            for (int c = lane; c < NC(); c += kTS) {
                const T* f = st_u(k) + nq + 3 * c;
                const T* df = zk + nq + 3 * c;
                const T f0 = f[0] + df[0], f1 = f[1] + df[1], f2 = f[2] + df[2];
                T g0 = 0, g1 = 0, g2 = 0;
                for (int which = 0; which < 5; ++which) {
                    const int r = nbx + 5 * c + which;
                    const V3<T> a = fric_coeff(c, which);
                    const T val = a.x * f0 + a.y * f1 + a.z * f2;
                    const Quad q = recs(k)[2 * r];
                    T corr = T(0);
                    if (corrector) {
                        const Quad dd = recs(k)[2 * r + 1];
                        corr = dd.v[0] * dd.v[2];
                    }
                    const T cf = side_coef(q.v[0], q.v[2], val, eps, mu_target, corr);
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
            const T eps = row_eps(3);
            T* crow = sV + 4 * nz;
            for (int i = lane; i < NOBS(); i += kTS) {
                const int r = nbx + NFRIC() + i;
                const T* J = ws + oLJO() + (k * NOBS() + i) * OBSW();
                T val = ws[oLHO() + k * NOBS() + i];
                for (int j = 0; j < OBSW(); ++j) val += J[j] * zk[nu + j];
                const Quad q = recs(k)[2 * r];
                T corr = T(0);
                if (corrector) {
                    const Quad dd = recs(k)[2 * r + 1];
                    corr = dd.v[0] * dd.v[2];
                }
                crow[i] = side_coef(q.v[0], q.v[2], val, eps, mu_target, corr);
            }
            tsync();
            if (lane < OBSW()) {
                T acc = 0;
                for (int i = 0; i < NOBS(); ++i) acc += crow[i] * ws[oLJO() + (k * NOBS() + i) * OBSW() + lane];
                vec[nu + lane] += acc;
            }
            tsync();
        }
    }

    // Factor-block staging: the block of the NEXT stage is fetched with cp.async (LDGSTS, generic proxy — no
    // proxy fence against the plain stores of the factor pass) into the other half of the idle stage-matrix
    // buffer while the current stage is processed.
    static constexpr int kFStrideStatic = ((D::nz * (D::nu | 1) + 3) & ~3) > ((D::nu * (D::nz + 1) + 3) & ~3)
                                              ? ((D::nz * (D::nu | 1) + 3) & ~3)
                                              : ((D::nu * (D::nz + 1) + 3) & ~3);
    static constexpr bool kFacDouble = D::kStatic && 2 * kFStrideStatic <= D::nz * (D::nz | 1);
    static_assert(!kStageTT || kFacDouble, "side-record staging shares the cp.async group schedule of the factor ring");
    __device__ __forceinline__ void fac_issue(int k, int buf) {
        if (k < 0 || k >= NN()) return;
        constexpr int V = 16 / sizeof(T);
        const T* src = ws + oFAC() + k * FSTRIDE();
        T* dst = sM + buf * FSTRIDE();
        for (int i = lane; i < FSTRIDE() / V; i += kTS) cp_async16(dst + i * V, src + i * V);
    }
    // equality rows of stage k (k < N) straight into sSA
    static constexpr bool kStageEQ = kStageTT && D::neq > 0 && (D::neq * D::nz) % 4 == 0;
    __device__ __forceinline__ void eq_issue(int k) const {
        if constexpr (kStageEQ) {
            if (k < 0 || k >= NN()) return;
            constexpr int V = 16 / sizeof(T);
            const T* src = ws + oLCT() + k * NEQ() * NZ();
            for (int i = lane; i < NEQ() * NZ() / V; i += kTS) cp_async16(sSA + i * V, src + i * V);
        }
    }
    // 16-byte vectorised copy global -> shared (both 16-byte aligned; n in elements)
    __device__ __forceinline__ void copy_block(T* __restrict__ dst, const T* __restrict__ src, int n) const {
        constexpr int V = 16 / sizeof(T);
        const int nv = n / V;
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(dst);
        for (int i = lane; i < nv; i += kTS) d4[i] = s4[i];
        for (int i = nv * V + lane; i < n; i += kTS) dst[i] = src[i];
    }

    // ---------------------------------------------------------------- fused IPM passes
    // Pass A (backward): per stage load the equality rows once, form the predictor gradient
    // (sigma = 0), keep it in GP for the corrector, build and factor the stage matrix and do the
    // backward vector step with the factor still in shared memory.
    __device__ bool pass_factor_predict() {
        const int nu = NU(), nx = NX(), nz = NZ(), ld = LDM(), ldf = LDF();
        T* vec = sV;
        bool ok = true;
        for (int i = lane; i < nx; i += kTS) sPv[i] = T(0);
        tsync();
        if constexpr (kStageTT) {
            tt_issue(NN());
            sm_issue(NN(), false, true);
            cp_commit();
        }
        for (int k = NN(); k >= 0; --k) {
            long long f0 = clock64();
            if constexpr (kStageTT) cp_wait<0>();           // side records (and equality rows) of stage k are staged
            stage_gradient(k, false, T(0), vec, kStageEQ && k < NN());  // equality rows of stage k in sSA afterwards
            T* GPk = ws + oLAM() + k * nz;
            for (int i = lane; i < nz; i += kTS) GPk[i] = vec[i];
#ifdef UB_DEBUG_NAN
            {
                T bad = 0, badr = 0;
                for (int i = lane; i < nz; i += kTS) bad += (vec[i] == vec[i]) ? T(0) : T(1);
                for (int r = lane; r < NROW(); r += kTS) {
                    const Quad q = recs(k)[2 * r];
                    for (int c = 0; c < 4; ++c) badr += (q.v[c] == q.v[c] && fabs(q.v[c]) < T(1e30)) ? T(0) : T(1);
                    if (row_valid(k, row_family(r)) && (!(q.v[0] > T(0)) || !(q.v[2] > T(0)))) badr += T(100);
                }
                bad = tsum(bad);
                badr = tsum(badr);
                if (nan_reason == 0 && badr > T(0)) nan_reason = 20000 + 100 * k + int(badr > T(99));
                if (nan_reason == 0 && bad > T(0)) nan_reason = 10000 + 100 * k;
            }
#endif
            long long f1 = clock64();
            t_g += f1 - f0;
            constexpr bool kAssignDyn = D::kStatic;         // dynamics term first (assigned), the rest added on top
            if (kAssignDyn && k < NN()) add_dynamics_hessian<true>();
            build_stage_matrix(k, true, kAssignDyn && k < NN());
            if constexpr (kStageTT) {                       // next stage's records / rows arrive during the factorisation
                tt_issue(k - 1);
                eq_issue(k - 1);
                sm_issue(k - 1, false, true);
                cp_commit();
            }
            long long f2 = clock64();
            t_f1 += f2 - f1;
            if (k < NN()) {
                if (!kAssignDyn) add_dynamics_hessian<false>();
                add_dynamics_gradient(vec);                  // uses p_{k+1} in sPv
                long long f3 = clock64();
                t_f2 += f3 - f2;
                T* F = ws + oFAC() + k * FSTRIDE();
                T* Wk = ws + oWF() + k * nu;
                if constexpr (kBlocked) {
                    // factor, forward substitution (w in vec[0, nu)), p -> sPv, P -> sP, factor block -> workspace
                    ok &= stage_factor_blocked<D::nu, D::nx>(vec, F);
                    t_f3 += clock64() - f3;
                    for (int j = lane; j < nu; j += kTS) Wk[j] = vec[j];
                } else {
                    ok &= stage_cholesky();
                    t_f3 += clock64() - f3;
                    for (int j = 0; j < nu; ++j) {  // forward substitution, column oriented
                        const T wj = vec[j] * sM[j * ld + j];
                        tsync();
                        if (lane == 0) vec[j] = wj;
                        for (int i = j + 1 + lane; i < nu; i += kTS) vec[i] -= sM[i * ld + j] * wj;
                        tsync();
                    }
                    for (int j = lane; j < nu; j += kTS) Wk[j] = vec[j];
                    // p = m_x - Y' w
                    for (int i = lane; i < nx; i += kTS) {
                        T acc = vec[nu + i];
                        const T* Mr = sM + (nu + i) * ld;
                        for (int j = 0; j < nu; ++j) acc -= Mr[j] * vec[j];
                        sPv[i] = acc;
                    }
                    // factor block [L; Y] -> workspace for the forward / corrector passes
                    for (int idx = lane; idx < nz * nu; idx += kTS) {
                        const int i = idx / nu, j = idx % nu;
                        T v = T(0);
                        if (j <= i) v = sM[i * ld + j];
                        F[i * ldf + j] = v;
                    }
                }
            } else {
                for (int i = lane; i < nx; i += kTS) sPv[i] = vec[nu + i];
                copy_cost_to_go();
            }
            tsync();
        }
        return ok;
    }

    // Pass C (backward, corrector): gradient = stored predictor gradient + the side terms that change
    // with the centring target and the second-order correction; backward vector step with stored factors.
    __device__ void pass_backward_corrector(T target_mu) {
        const int nq = NQ(), nu = NU(), nx = NX(), nz = NZ(), ldf = LDF();
        T* vec = sV;
        for (int i = lane; i < nx; i += kTS) sPv[i] = T(0);
        tsync();
        const int nbx = NBOXU() + nx;
        // cp.async group schedule: TT(k) is committed before FAC(k); every wait leaves exactly one younger group
        // in flight (none at the terminal stage)
        if constexpr (kFacDouble) {
            tt_issue(NN());
            sm_issue(NN(), true, false);
            cp_commit();
        }
        for (int k = NN(); k >= 0; --k) {
            if constexpr (kFacDouble) {
                if (k == NN()) cp_wait<0>();
                else cp_wait<1>();
            }
            const T* GPk = st_gp(k);
            for (int i = lane; i < nz; i += kTS) vec[i] = GPk[i];
            tsync();
            // (corr - target) / (t + eps lam) per side
            for (int r = lane; r < nbx; r += kTS) {
                const int fam = r < NBOXU() ? 0 : 1;
                if (!row_valid(k, fam)) continue;
                const int m = fam == 0 ? r : nu + (r - NBOXU());
                const T eps = row_eps(fam);
                const Quad q = recs(k)[2 * r];
                const Quad dd = recs(k)[2 * r + 1];
                vec[m] += fdiv(dd.v[0] * dd.v[2] - target_mu, q.v[0] + eps * q.v[2]) -
                          fdiv(dd.v[1] * dd.v[3] - target_mu, q.v[1] + eps * q.v[3]);
            }
            tsync();
            if (NFRIC() > 0 && k < NN()) {
                const T eps = row_eps(2);
                for (int c = lane; c < NC(); c += kTS) {
                    T g0 = 0, g1 = 0, g2 = 0;
                    for (int which = 0; which < 5; ++which) {
                        const int r = nbx + 5 * c + which;
                        const V3<T> a = fric_coeff(c, which);
                        const Quad q = recs(k)[2 * r];
                        const Quad dd = recs(k)[2 * r + 1];
                        const T cf = fdiv(dd.v[0] * dd.v[2] - target_mu, q.v[0] + eps * q.v[2]);
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
                const T eps = row_eps(3);
                T* crow = sV + 4 * nz;
                for (int i = lane; i < NOBS(); i += kTS) {
                    const int r = nbx + NFRIC() + i;
                    const Quad q = recs(k)[2 * r];
                    const Quad dd = recs(k)[2 * r + 1];
                    crow[i] = fdiv(dd.v[0] * dd.v[2] - target_mu, q.v[0] + eps * q.v[2]);
                }
                tsync();
                if (lane < OBSW()) {
                    T acc = 0;
                    for (int i = 0; i < NOBS(); ++i) acc += crow[i] * ws[oLJO() + (k * NOBS() + i) * OBSW() + lane];
                    vec[nu + lane] += acc;
                }
                tsync();
            }
            if constexpr (kFacDouble) {
                tsync();
                tt_issue(k - 1);
                sm_issue(k - 1, true, false);
                cp_commit();
            }
            if (k == NN()) {
                for (int i = lane; i < nx; i += kTS) sPv[i] = vec[nu + i];
                if constexpr (kFacDouble) {
                    fac_issue(k - 1, (k - 1) & 1);
                    cp_commit();
                }
                tsync();
                continue;
            }
            add_dynamics_gradient(vec);
            const T* F = sM;
            if constexpr (kFacDouble) {
                cp_wait<1>();
                fac_issue(k - 1, (k - 1) & 1);
                cp_commit();
                F = sM + (k & 1) * FSTRIDE();
            } else {
                copy_block(sM, ws + oFAC() + k * FSTRIDE(), nz * ldf);
                tsync();
            }
            T* Wk = ws + oWF() + k * nu;
            for (int j = 0; j < nu; ++j) {
                const T wj = vec[j] * F[fidx(j, j)];
                tsync();
                if (lane == 0) vec[j] = wj;
                for (int i = j + 1 + lane; i < nu; i += kTS) vec[i] -= F[fidx(i, j)] * wj;
                tsync();
            }
            for (int j = lane; j < nu; j += kTS) Wk[j] = vec[j];
        
            for (int i = lane; i < nx; i += kTS) {
                T acc = vec[nu + i];
                for (int j = 0; j < nu; ++j) acc -= F[fidx(nu + i, j)] * vec[j];
                sPv[i] = acc;
            }
            tsync();
        }
        if constexpr (kFacDouble) cp_wait<0>();
    }

    // Side steps of the rows of ONE stage for the stage direction d = [du; dx] (shared memory):
    // d lambda, d t per side, and the running maximum feasible step.
    __device__ __forceinline__ void stage_side_steps(int k, const T* d, bool corrector, T target_mu, T& amax) {
        const T* zk = st_z(k);
        const T cm = corrector ? T(1) : T(0);
        for (int r = lane; r < NROW(); r += kTS) {
            const int fam = row_family(r);
            if (!row_valid(k, fam)) continue;
            T lb, ub;
            const T val = row_value(k, r, fam, zk, st_x(k), st_u(k), &lb, &ub);
            const T adz = row_dot(k, r, fam, d);
            const T eps = row_eps(fam);
            const Quad q = recs(k)[2 * r];
            Quad dd = recs(k)[2 * r + 1];
            const int nsd = fam >= 2 ? 1 : 2;
            for (int sd = 0; sd < nsd; ++sd) {
                const T t = q.v[sd], lam = q.v[2 + sd];
                const T sg = sd == 0 ? T(1) : T(-1);
                const T dist = sd == 0 ? val - lb : ub - val;
                const T rd = dist + eps * lam - t;
                const T rc = t * lam - target_mu + cm * dd.v[sd] * dd.v[2 + sd];
                const T den = t + eps * lam;
                const T iden = fdiv(T(1), den);
                const T dl = -(rc + lam * rd) * iden - (lam * iden) * sg * adz;
                const T dtt = sg * adz + eps * dl + rd;
                dd.v[sd] = dtt;
                dd.v[2 + sd] = dl;
                if (dtt < T(0)) amax = min(amax, fdiv(-t, dtt));
                if (dl < T(0)) amax = min(amax, fdiv(-lam, dl));
            }
            *side_dd(k, r) = dd;
        }
    }

    // Passes B / D (forward): direction from the stored factors and w, written to DZ, with the side
    // steps of every stage fused in.  Returns the largest feasible step in (0, 1].
    __device__ T pass_forward(bool corrector, T target_mu) {
        const int nq = NQ(), nu = NU(), nx = NX(), nz = NZ(), ldf = LDF();
        T* dxn = sV + nz;       // [nx] next state direction
        T* dst = sV + 2 * nz;   // [nz] stage direction [du; dx]
        T* du = dst;
        T* dx = dst + nu;
        T amax = T(1);
        for (int i = lane; i < nz; i += kTS) dst[i] = T(0);
        tsync();
        // cp.async group schedule: FAC(k) is committed before TT(k); every wait leaves exactly one younger group
        // in flight (none for the records of the terminal stage)
        if constexpr (kFacDouble) {
            fac_issue(0, 0);
            w_issue(0);
            cp_commit();
            tt_issue(0);
            sm_issue(0, false, false);
            cp_commit();
        }
        for (int k = 0; k <= NN(); ++k) {
            if (k < NN()) {
                const T* F = sM;
                const T* Wk = ws + oWF() + k * nu;
                T wreg = T(0);
                if constexpr (kFacDouble) {
                    cp_wait<1>();
                    if constexpr (kStageTT) {   // w_k leaves its (single) staging slot before w_{k+1} is requested
                        static_assert(D::nu <= kTS, "one register per lane holds the staged w");
                        if (lane < nu) wreg = st_w(k)[lane];
                        tsync();
                    }
                    fac_issue(k + 1, (k + 1) & 1);
                    w_issue(k + 1);
                    cp_commit();
                    F = sM + (k & 1) * FSTRIDE();
                } else {
                    copy_block(sM, ws + oFAC() + k * FSTRIDE(), nz * ldf);
                    tsync();
                }
                // s = w + Y dx
                for (int j = lane; j < nu; j += kTS) {
                    T acc = kStageTT ? wreg : Wk[j];
                    for (int i = 0; i < nx; ++i) acc += F[fidx(nu + i, j)] * dx[i];
                    du[j] = acc;
                }
                tsync();
                for (int j = nu - 1; j >= 0; --j) {
                    const T uj = -du[j] * F[fidx(j, j)];
                    tsync();
                    for (int i = lane; i < j; i += kTS) du[i] += F[fidx(j, i)] * uj;
                    if (lane == 0) du[j] = uj;
                    tsync();
                }
            
            } else {
                for (int j = lane; j < nu; j += kTS) du[j] = T(0);
                tsync();
            }
            T* Dk = DZk(k);
            for (int i = lane; i < nz; i += kTS) Dk[i] = dst[i];
            if constexpr (kFacDouble) {
                if (k == NN()) cp_wait<0>();
                else cp_wait<1>();
            }
            stage_side_steps(k, dst, corrector, target_mu, amax);
            if constexpr (kFacDouble) {
                tsync();
                tt_issue(k + 1);
                sm_issue(k + 1, false, false);
                cp_commit();
            }
            if (k < NN()) {
                if (lane < nq) {
                    const T dt = C.dt;
                    const T q = dx[lane], v = dx[nq + lane], a = dx[2 * nq + lane], j = du[lane];
                    dxn[lane] = q + dt * v + T(0.5) * dt * dt * a + dt * dt * dt / T(6) * j;
                    dxn[nq + lane] = v + dt * a + T(0.5) * dt * dt * j;
                    dxn[2 * nq + lane] = a + dt * j;
                }
                tsync();
                for (int i = lane; i < nx; i += kTS) dx[i] = dxn[i];
                tsync();
            }
        }
        if constexpr (kFacDouble) cp_wait<0>();
        return tmin(amax);
    }

    // Interior-point QP solve around the current (X, U).  Leaves the step in Z
    // and the factors of the last iteration in FAC.  Returns iterations used;
    // *converged, *decr as in orc::solve_qp_ipm.
    __device__ int solve_qp(bool* converged, T* decr, bool* finite) {
        const int nq = NQ(), nu = NU(), nx = NX(), nz = NZ(), N = NN();
        *converged = false;
        *finite = true;
        // dynamics-feasible start: du = 0, dx_0 = 0, dx_{k+1} = A dx_k + gap_k
        for (int idx = lane; idx < (N + 1) * nz; idx += kTS) ws[oZ() + idx] = T(0);
        tsync();
        if (lane < nq) {
            const T dt = C.dt;
            T q = 0, v = 0, a = 0;
            for (int k = 0; k < N; ++k) {
                const T* gap = ws + oGAP() + k * nx;
                const T qn = q + dt * v + T(0.5) * dt * dt * a + gap[lane];
                const T vn = v + dt * a + gap[nq + lane];
                const T an = a + gap[2 * nq + lane];
                q = qn; v = vn; a = an;
                T* zn = Zk(k + 1);
                zn[nu + lane] = q;
                zn[nu + nq + lane] = v;
                zn[nu + 2 * nq + lane] = a;
            }
        }
        tsync();
        init_eq_weights();
        // slack / multiplier initialisation
        int nsides_l = 0;
        for (int k = 0; k <= N; ++k) {
            const T* zk = Zk(k);
            for (int r = lane; r < NROW(); r += kTS) {
                const int fam = row_family(r);
                Quad q, dd;
                q.v[0] = q.v[1] = T(1);
                q.v[2] = q.v[3] = T(0);
                dd.v[0] = dd.v[1] = dd.v[2] = dd.v[3] = T(0);
                if (row_valid(k, fam)) {
                    T lb, ub;
                    const T val = row_value(k, r, fam, zk, X + k * NX(), U + k * NU(), &lb, &ub);
                    q.v[0] = max(val - lb, C.thr0);
                    q.v[2] = C.mu0 / q.v[0];
                    ++nsides_l;
                    if (fam < 2) {
                        q.v[1] = max(ub - val, C.thr0);
                        q.v[3] = C.mu0 / q.v[1];
                        ++nsides_l;
                    }
                }
                *side_tl(k, r) = q;
                *side_dd(k, r) = dd;
            }
        }
        const int nsides = int(tsum(T(nsides_l)) + T(0.5));
        tsync();
        T last_alpha = T(0), last_step = tinf<T>();
        int iters = 0;
        // initial residual summary (afterwards mu comes from the update pass and the slack residual
        // contracts by exactly (1 - alpha) per Newton step, the rows being linear)
        T mu = 0, rdmax = 0;
        for (int k = 0; k <= N; ++k) {
            const T* zk = Zk(k);
            for (int r = lane; r < NROW(); r += kTS) {
                const int fam = row_family(r);
                if (!row_valid(k, fam)) continue;
                T lb, ub;
                const T val = row_value(k, r, fam, zk, X + k * NX(), U + k * NU(), &lb, &ub);
                const T eps = row_eps(fam);
                const Quad q = *side_tl(k, r);
                rdmax = max(rdmax, fabs(val - lb + eps * q.v[2] - q.v[0]));
                mu += q.v[0] * q.v[2];
                if (fam < 2) {
                    rdmax = max(rdmax, fabs(ub - val + eps * q.v[3] - q.v[1]));
                    mu += q.v[1] * q.v[3];
                }
            }
        }
        mu = nsides > 0 ? tsum(mu) / T(nsides) : T(0);
        rdmax = tmax(rdmax);
        T pinf = T(0);
        auto eq_infeasibility = [&]() {
            T pv = T(0);
            if (!C.soft_poly)
                for (int k = 0; k <= N; ++k) {
                    if (neq_of(k) == 0) continue;
                    load_eq_rows(k);
                    const T* zk = Zk(k);
                    for (int i = lane; i < neq_of(k); i += kTS)
                        if (rho_eq(k)[i] > T(0)) pv = max(pv, fabs(eq_value(k, i, zk)));
                    tsync();
                }
            return tmax(pv);
        };
        pinf = eq_infeasibility();
        for (int it = 0; it < C.qp_iter_max; ++it) {
            const long long c_it = clock64();
            if (it > 0 && mu <= T(2) * C.mu_target && rdmax <= C.qp_tol && last_alpha >= T(0.5) &&
                (pinf <= C.qp_tol || last_step <= C.qp_tol)) {
                *converged = true;
                break;
            }
            iters = it + 1;
            if (!pass_factor_predict()) {
                *finite = false;
                if (nan_reason == 0) nan_reason = 1;
            }
            long long c2 = clock64();
            t_fac += c2 - c_it;
            T target_mu = C.mu_target;
            T alpha = T(1);
            // predictor (pass 0) and corrector (pass 1) share ONE inlined copy of the forward pass
            T a_fwd = T(1);
#pragma unroll 1
            for (int pass = 0; pass < (nsides > 0 ? 2 : 1); ++pass) {
                if (pass == 1) {
                    // mean complementarity after the affine step -> centring target (Mehrotra)
                    const T a_aff = a_fwd;
                    T acc = 0;
                    for (int idx = lane; idx < (N + 1) * NROW(); idx += kTS) {
                        const int k = idx / NROW(), r = idx % NROW();
                        const int fam = row_family(r);
                        if (!row_valid(k, fam)) continue;
                        const Quad* rec = reinterpret_cast<const Quad*>(ws + oTT()) + 2 * idx;
                        const Quad q = rec[0], dd = rec[1];
                        acc += (q.v[0] + a_aff * dd.v[0]) * (q.v[2] + a_aff * dd.v[2]);
                        if (fam < 2) acc += (q.v[1] + a_aff * dd.v[1]) * (q.v[3] + a_aff * dd.v[3]);
                    }
                    const T mu_aff = tsum(acc) / T(nsides);
                    const T ratio = mu_aff / mu;
                    target_mu = max(ratio * ratio * ratio * mu, C.mu_target);
                    long long c3 = clock64();
                    t_swp += c3 - c2;
                    c2 = c3;
                    pass_backward_corrector(target_mu);
                }
                a_fwd = pass_forward(pass == 1, pass == 1 ? target_mu : T(0));
            }
            if (nsides > 0) alpha = min(T(1), T(0.995) * a_fwd);
            t_side += clock64() - c2;
            // update z, t, lambda; new mean complementarity
            T stepmax = 0, musum = 0;
            for (int idx = lane; idx < (N + 1) * nz; idx += kTS) {
                const T d = ws[oDZ() + idx];
                ws[oZ() + idx] += alpha * d;
                stepmax = max(stepmax, fabs(alpha * d));
            }
            for (int idx = lane; idx < (N + 1) * NROW(); idx += kTS) {
                const int k = idx / NROW(), r = idx % NROW();
                const int fam = row_family(r);
                Quad* rec = reinterpret_cast<Quad*>(ws + oTT()) + 2 * idx;
                Quad q = rec[0];
                const Quad dd = rec[1];
#pragma unroll
                for (int c = 0; c < 4; ++c) q.v[c] += alpha * dd.v[c];
                rec[0] = q;
                if (row_valid(k, fam)) {
                    musum += q.v[0] * q.v[2];
                    if (fam < 2) musum += q.v[1] * q.v[3];
                }
            }
            tsync();
            mu = nsides > 0 ? tsum(musum) / T(nsides) : T(0);
            rdmax *= (T(1) - alpha);
            if (!C.soft_poly) {
                // multiplier update of the hard equality rows and their infeasibility at the new iterate, one visit
                T pv = T(0);
                for (int k = 0; k <= N; ++k) {
                    if (neq_of(k) == 0) continue;
                    load_eq_rows(k);
                    const T* zk = Zk(k);
                    for (int i = lane; i < neq_of(k); i += kTS) {
                        const T rho = rho_eq(k)[i];
                        if (rho > T(0)) {
                            const T e = eq_value(k, i, zk);
                            y_eq(k)[i] += rho * e;
                            pv = max(pv, fabs(e));
                        }
                    }
                    tsync();
                }
                pinf = tmax(pv);
            }
            last_alpha = alpha;
            last_step = tmax(stepmax);
            if (!(last_step < tinf<T>())) {
                *finite = false;
                if (nan_reason == 0) nan_reason = !(mu < tinf<T>()) ? 4 : (!(alpha <= T(1)) ? 5 : 2);
            }
            if (!*finite) break;
        }
        *decr = last_step;
        return iters;
    }

    // Feedback gains K_k = -Huu^{-1} Hux from the stored factors (optional output).
    __device__ void write_gains(T* Kout, int stages) {
        const int nu = NU(), nx = NX();
        T* F = sM;
        T* col = sV;
        for (int k = 0; k < stages; ++k) {
            const T* Fg = ws + oFAC() + k * FSTRIDE();
            for (int idx = lane; idx < FSTRIDE(); idx += kTS) F[idx] = Fg[idx];
            tsync();
            for (int xcol = 0; xcol < nx; ++xcol) {
                for (int j = lane; j < nu; j += kTS) col[j] = F[fidx(nu + xcol, j)];
                tsync();
                for (int j = nu - 1; j >= 0; --j) {
                    const T uj = -col[j] * F[fidx(j, j)];
                    tsync();
                    for (int i = lane; i < j; i += kTS) col[i] += F[fidx(j, i)] * uj;
                    if (lane == 0) col[j] = uj;
                    tsync();
                }
            
                for (int j = lane; j < nu; j += kTS) Kout[(k * nu + j) * nx + xcol] = col[j];
                tsync();
            }
        }
    }

    // --------------------------------------------------------------- solve
    __device__ void run(const BatchArgs<T>& A, int b) {
        const int nq = NQ(), nx = NX(), nu = NU(), N = NN(), nz = NZ();
        t_lin = t_fac = t_swp = t_side = t_ls = t_f1 = t_f2 = t_f3 = t_g = 0;
        nan_reason = 0;
        // initial guess: DefaultInitializer = zero input, state held
        // (controller_interface.cpp:385-386); x_0 is always the observation
        {
            const int nxt = A.nxt, nxo = NXO();
            const T* x0 = A.x0 + size_t(b) * nxt;
            T* XOw = ws + oXO();
            if (!A.warm) {
                for (int idx = lane; idx < (N + 1) * nx; idx += kTS) X[idx] = x0[idx % nx];
                for (int idx = lane; idx < N * nu; idx += kTS) U[idx] = T(0);
                for (int idx = lane; idx < (N + 1) * nxo; idx += kTS) XOw[idx] = x0[nx + idx % nxo];   // state held
            } else {
                const T* Xin = A.Xin + size_t(b) * (N + 1) * nxt;
                const T* Uin = A.Uin + size_t(b) * N * nu;
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
            const T* tg = A.target + size_t(b) * (N + 1) * 3;
            T* tgl = ws + oTG();
            for (int idx = lane; idx < (N + 1) * 3; idx += kTS) tgl[idx] = tg[idx];
            const T* bd = A.body ? A.body + size_t(b) * NB() * UB_BODY_PARAMS : &P.body[0][0];
            T* bdl = ws + oBD();
            for (int idx = lane; idx < NB() * UB_BODY_PARAMS; idx += kTS) bdl[idx] = bd[idx];
        }
        tsync();
        if (NEQ() > 0) build_Df();
        Perf<T> base = performance(X, U);
        int status = UB_STATUS_CONVERGED, qp_iters = 0, sqp_done = 0;
        T alpha = 0, qp_res = 0;
        T* Xn = ws + oXN();
        T* Un = ws + oUN();
        for (int it = 0; it < max(1, C.sqp_iters); ++it) {
            ++sqp_done;
            long long c0 = clock64();
            if (NXO() > 0) {
                // Newton step of the obstacle states = exact constant-acceleration rollout of the observation - iterate
                const T* o0 = ws + oXO();
                for (int idx = lane; idx < (N + 1) * P.ndyn * 3; idx += kTS) {
                    const int k = idx / (P.ndyn * 3), j = (idx / 3) % P.ndyn, c = idx % 3;
                    const T t = C.dt * T(k);
                    const T p0 = o0[j * 9 + c], v0 = o0[j * 9 + 3 + c], a0 = o0[j * 9 + 6 + c];
                    const int o = (k * P.ndyn + j) * 9;
                    ws[oDXO() + o + c] = p0 + t * v0 + T(0.5) * t * t * a0 - ws[oXO() + o + c];
                    ws[oDXO() + o + 3 + c] = v0 + t * a0 - ws[oXO() + o + 3 + c];
                    ws[oDXO() + o + 6 + c] = a0 - ws[oXO() + o + 6 + c];
                }
                tsync();
            }
            linearize();
            t_lin += clock64() - c0;
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
            T desc = 0;
            for (int k = 0; k < N; ++k) {
                const T* zk = Zk(k);
                const T* x = X + k * nx;
                const T* u = U + k * nu;
                const T* Jp = ws + oLJP() + k * 3 * nq;
                for (int i = lane; i < nz; i += kTS) {
                    T g;
                    if (i < nq) g = C.dt * P.Rd[i] * u[i];
                    else if (i < nu) g = C.dt * C.fw * u[i];
                    else {
                        const int xi = i - nu;
                        g = C.dt * P.Qd[xi] * (x[xi] - P.xd[xi]);
                        if (xi < nq)
                            for (int c = 0; c < 3; ++c)
                                g += C.dt * C.Wd[c] * Jp[c * nq + xi] * (ws[oLR() + 3 * k + c] - target[3 * k + c]);
                        if (IALIGN())
                            for (int r = 0; r < 2; ++r)
                                g += C.dt * P.ia_w * ws[oLJA() + (2 * k + r) * nx + xi] * ws[oLIA() + 2 * k + r];
                    }
                    desc += g * zk[i];
                }
            }
            desc = tsum(desc);
            const long long c_ls = clock64();
            // filter line search (ocs2 FilterLinesearch [EXT])
            const T vb = base.violation();
            bool accepted = false;
            Perf<T> pn = base;
            alpha = T(1);
            while (alpha >= C.alpha_min) {
                for (int idx = lane; idx < (N + 1) * nx; idx += kTS) {
                    const int k = idx / nx, i = idx % nx;
                    Xn[idx] = X[idx] + alpha * ws[oZ() + k * nz + nu + i];
                }
                for (int idx = lane; idx < N * nu; idx += kTS) {
                    const int k = idx / nu, i = idx % nu;
                    Un[idx] = U[idx] + alpha * ws[oZ() + k * nz + i];
                }
                tsync();
                pn = performance(Xn, Un, alpha);
                const T vn = pn.violation();
                if (vn > C.g_max) accepted = false;
                else if (vn < C.g_min) {
                    if (vb < C.g_min && desc < T(0)) accepted = pn.cost < base.cost + C.armijo * alpha * desc;
                    else accepted = true;
                } else {
                    accepted = (vn < (T(1) - C.gamma_c) * vb) || (pn.cost < base.cost - C.gamma_c * vb);
                }
                if (accepted) break;
                alpha *= C.alpha_decay;
            }
            t_ls += clock64() - c_ls;
            if (!accepted) {
                status = UB_STATUS_LS_FAILED;
                alpha = T(0);
                break;
            }
            T dxn = 0, dun = 0;
            for (int idx = lane; idx < (N + 1) * nx; idx += kTS) {
                const T d = Xn[idx] - X[idx];
                dxn += d * d;
                X[idx] = Xn[idx];
            }
            for (int idx = lane; idx < N * nu; idx += kTS) {
                const T d = Un[idx] - U[idx];
                dun += d * d;
                U[idx] = Un[idx];
            }
            for (int idx = lane; idx < (N + 1) * NXO(); idx += kTS) ws[oXO() + idx] += alpha * ws[oDXO() + idx];
            dxn = sqrt(tsum(dxn));
            dun = sqrt(tsum(dun));
            tsync();
            const T dcost = fabs(pn.cost - base.cost);
            base = pn;
            if ((dxn < C.delta_tol && dun < C.delta_tol) || (dcost < C.cost_tol && base.violation() < C.g_min)) break;
        }
        if (A.K != nullptr && A.stop_after == 0 && status != UB_STATUS_NAN)
            write_gains(A.K + size_t(b) * A.gain_stages * nu * nx, A.gain_stages);
        // solution out + NaN guard
        T bad = 0;
        {
            const int nxt = A.nxt, nxo = NXO();
            T* Xo = A.X + size_t(b) * (N + 1) * nxt;
            T* Uo = A.U + size_t(b) * N * nu;
            for (int idx = lane; idx < (N + 1) * nx; idx += kTS) {
                const T v = X[idx];
                bad += isfinite(v) ? T(0) : T(1);
                Xo[(idx / nx) * nxt + idx % nx] = v;
            }
            for (int idx = lane; idx < (N + 1) * nxo; idx += kTS) Xo[(idx / nxo) * nxt + nx + idx % nxo] = ws[oXO() + idx];
            for (int idx = lane; idx < N * nu; idx += kTS) Uo[idx] = U[idx];
        }
        bad = tsum(bad);
        if (bad > T(0)) {
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
                T* s = A.stats + size_t(b) * UB_STATS;
                s[0] = T(qp_iters);
                s[1] = base.cost;
                s[2] = base.violation();
                s[3] = status == UB_STATUS_NAN ? T(nan_reason) : alpha;   // NaN status: where it first appeared
                s[4] = qp_res;
                s[5] = base.max_eq;
                s[6] = base.min_margin;
                s[7] = T(sqp_done);
                if (A.stop_after == 9) {  // profile mode: phase cycle counters replace stats[1..7]
                    s[1] = T(t_g);
                    s[2] = T(t_f1);
                    s[3] = T(t_f2);
                    s[4] = T(t_f3);
                    s[5] = T(t_fac);
                    s[6] = T(t_swp);
                    s[7] = T(t_side);
                }
                __threadfence_system();
            }
            *reinterpret_cast<volatile int32_t*>(A.status + b) = status;
        }
    }
};

template <typename T, typename D, int TW = 1>
__global__ void __launch_bounds__(256, 2) solve_batch_kernel(const __grid_constant__ DevProblem<T> Pc, const DevProblem<T>* __restrict__ Pg,
                                                          Layout L, BatchArgs<T> A, int teams_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // CTA-shared copy of the problem constants
    DevProblem<T>* Ps = reinterpret_cast<DevProblem<T>*>(smem_raw);
    {
        const int nwords = sizeof(DevProblem<T>) / 4;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Pg);
        uint32_t* dst = reinterpret_cast<uint32_t*>(Ps);
        for (int i = threadIdx.x; i < nwords; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    constexpr int kTeam = TW * WARP;
    const int team = int(threadIdx.x) / kTeam, lane = threadIdx.x % kTeam;   // "lane" = thread of the instance team
    const int slot = blockIdx.x * teams_per_cta + team;
    if (slot >= A.n_slots) return;   // whole teams leave together (their named barrier is theirs alone)
    constexpr size_t off = (sizeof(DevProblem<T>) + 15) / 16 * 16;
    Layout Lk = L;
    if constexpr (D::kStatic) Lk = D::template layout<T>();  // compile-time offsets (host passes the same numbers)
    // The per-team shared-memory offset and the per-slot workspace offset pass through an opaque move: the
    // compiler then keeps them in a register instead of re-deriving them from threadIdx / blockIdx at every use
    // (measured: that rematerialisation was ~8 % of all executed instructions).
    uint32_t sm_off = uint32_t(off + size_t(team) * Lk.s_total * sizeof(T));
    asm volatile("mov.u32 %0, %0;" : "+r"(sm_off));
    unsigned long long ws_off = (unsigned long long)(slot) * (unsigned long long)(Lk.total);
    asm volatile("mov.u64 %0, %0;" : "+l"(ws_off));
    T* sm = reinterpret_cast<T*>(smem_raw + sm_off);
    Solver<T, D, TW> S(*Ps, Pc, L, lane);
    S.ws = A.ws + ws_off;
    S.X = S.ws + Lk.XW;
    S.U = S.ws + Lk.UW;
    S.target = S.ws + Lk.TG;
    S.body = S.ws + Lk.BD;
    S.sM = sm + Lk.sM;
    S.sP = sm + Lk.sP;
    S.sPv = sm + Lk.sPv;
    S.sSA = sm + Lk.sSA;
    S.sV = sm + Lk.sV;
    S.sTT = sm + Lk.sTT;
    S.sSm = sm + Lk.sSm;
    S.sRed = sm + Lk.sRed;
    S.bar_id = 1 + team;             // barrier 0 is __syncthreads
    if (A.queue == nullptr) {   // static mode (test aid)
        S.run(A, slot);
        return;
    }
    for (;;) {
        int b = 0;
        if (lane == 0) b = atomicAdd(A.queue, 1);
        if constexpr (TW == 1) {
            b = __shfl_sync(FULL, b, 0);
        } else {   // broadcast through the team scratch
            S.tsync();
            if (lane == 0) *reinterpret_cast<volatile int*>(S.sRed + TW) = b;
            S.tsync();
            b = *reinterpret_cast<volatile int*>(S.sRed + TW);
        }
        if (b >= A.B) break;
        S.run(A, b);
        S.tsync();
    }
}

}  // namespace ub
