// TEST INFRASTRUCTURE (precision study, tools/precision_lab.py) — included by oracle.cpp inside namespace orc.
//
// The interior-point iteration of solve_qp_ipm with the linear algebra the round-2 kernels use
// (ub_solver.cuh, "reduced" stage): the contact forces are eliminated from every stage in RANGE-SPACE form before
// the Riccati recursion, which then runs on [jerk; state] only (nq + nx = 36 variables for every Thing
// configuration).  With the equality rows  E = [0 | Df | C],  weights R = diag(rho)  and the force block
// D = H_ff + barrier terms (diagonal, or 3 x 3 blocks per contact with friction pyramids):
//
//     S      = R^-1 + Df D^-1 Df'                       (6 nb x 6 nb, SPD; no penalty weight appears, only 1/rho)
//     M_xx  += C' S^-1 C                                 (what the penalty term leaves after the forces are gone)
//     m_x   += C' S^-1 (e + R^-1 y - Df D^-1 m_f)
//     lambda = S^-1 (e + R^-1 y - Df D^-1 m_f + C dx),   df = - D^-1 (m_f + Df' lambda)
//
// which is the exact block elimination of df from the Newton system of solve_qp_ipm (Woodbury on
// M_ff = D + Df' R Df), so <double, double, double, double> reproduces solve_qp_ipm to rounding.
// Types:  F  = Riccati matrices / factors / direction on [jerk; state],
//         R  = iterate, slack records, residuals, gradients,
//         ST = the force / equality block (D, S, its Cholesky factor, G = L^-1 C),
//         DT = storage type of the linearisation (rows, constants, bounds): rounds the QP data once.
template <typename F, typename R, typename ST, typename DT>
static void solve_qp_ipm_reduced(const ub_problem_desc_t& P, const Workspace& W, std::vector<std::vector<double>>& zout,
                                 double* info, int extra = 0, double mu_factor = 2.0) {
    int extra_left = extra;
    const Dims& D = W.D;
    const int nx = D.nx, nu = D.nu, N = D.N, nq = D.nq, nfc = D.nfc;
    // ORACLE_DATA_NOISE = relative perturbation of every datum of the linearisation (emulates Jacobians computed in fp32:
    // measured on the B200 at ~1e-6 absolute) — deterministic hash noise
    const double noise = std::getenv("ORACLE_DATA_NOISE") ? std::atof(std::getenv("ORACLE_DATA_NOISE")) : 0.0;
    unsigned long long hstate = 88172645463325252ull;
    auto rd_ = [&](double v) {
        if (noise > 0.0 && v != 0.0) {
            hstate ^= hstate << 13; hstate ^= hstate >> 7; hstate ^= hstate << 17;
            const double u = double(hstate >> 11) / double(1ull << 53) - 0.5;
            v += 2.0 * noise * u * std::max(1.0, std::fabs(v));
        }
        return R(DT(v));
    };
    enum Kind { JBOX, FBOX, XBOX, FRIC, XROW, EQF, EQX };
    struct RowR {
        Kind kind;
        int idx = -1;
        std::vector<R> a;
        R c = 0, lb = 0, ub = 0, rho = 0, lambda = 0;
        bool hard = false, on[2] = {false, false};
        R t[2] = {0, 0}, lam[2] = {0, 0}, dt[2] = {0, 0}, dl[2] = {0, 0};
    };
    struct StageR {
        int nz = 0, nu = 0;
        std::vector<R> H, g, b;
        std::vector<RowR> rows;
    };
    std::vector<StageR> st(N + 1);
    for (int k = 0; k <= N; ++k) {
        const Stage& s = W.st[k];
        StageR& q = st[k];
        q.nz = s.nz;
        q.nu = s.nu;
        q.H.resize(s.H.d.size());
        for (size_t i = 0; i < s.H.d.size(); ++i) q.H[i] = rd_(s.H.d[i]);
        q.g.resize(s.g.size());
        for (size_t i = 0; i < s.g.size(); ++i) q.g[i] = rd_(s.g[i]);
        q.b.resize(s.b.size());
        for (size_t i = 0; i < s.b.size(); ++i) q.b[i] = rd_(s.b[i]);
        for (const Row& r : s.rows) {
            RowR t;
            t.idx = r.idx;
            t.a.resize(r.a.size());
            for (size_t i = 0; i < r.a.size(); ++i) t.a[i] = rd_(r.a[i]);
            t.c = rd_(r.c);
            const bool eq = !(r.lb < r.ub);
            t.lb = rd_(std::isfinite(r.lb) ? r.lb : 0.0);
            t.ub = rd_(std::isfinite(r.ub) ? r.ub : 0.0);
            t.on[0] = !eq && std::isfinite(r.lb);
            t.on[1] = !eq && std::isfinite(r.ub);
            t.rho = R(r.rho);
            t.lambda = R(r.lambda);
            t.hard = r.hard;
            const int jn = q.nu > 0 ? nq : 0, fn = q.nu;   // [0, jn) jerk, [jn, fn) forces, [fn, nz) state
            if (r.idx >= 0) t.kind = r.idx < jn ? JBOX : (r.idx < fn ? FBOX : XBOX);
            else {
                bool tj = false, tf = false, tx = false;
                for (int i = 0; i < q.nz; ++i)
                    if (r.a[i] != 0.0) (i < jn ? tj : (i < fn ? tf : tx)) = true;
                if (tj) throw std::runtime_error("reduced stage: dense row on the jerk block");
                if (eq) t.kind = tf ? EQF : EQX;
                else if (tf && tx) throw std::runtime_error("reduced stage: inequality row on forces and state");
                else t.kind = tf ? FRIC : XROW;
            }
            q.rows.push_back(t);
        }
    }
    RiccatiT<F> ric;
    ric.nx = nx; ric.nu = nq; ric.N = N;
    ric.pivot_floor = F(P.reg_input);
    ric.A.assign(W.A.d.begin(), W.A.d.end());
    ric.B.assign(size_t(nx) * nq, F(0));
    for (int i = 0; i < nx; ++i)
        for (int j = 0; j < nq; ++j) ric.B[i * nq + j] = F(W.B(i, j));
    std::vector<std::vector<R>> z(N + 1);
    for (int k = 0; k <= N; ++k) z[k].assign(st[k].nz, R(0));
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < nx; ++i) {
            R v = st[k].b[i];
            for (int j = 0; j < nx; ++j) v += R(W.A(i, j)) * z[k][nu + j];
            z[k + 1][st[k + 1].nu + i] = v;
        }
    const R sgn[2] = {R(1), R(-1)};
    auto row_val = [](const RowR& r, const R* zz) {
        if (r.idx >= 0) return r.c + zz[r.idx];
        R v = r.c;
        for (size_t j = 0; j < r.a.size(); ++j) v += r.a[j] * zz[j];
        return v;
    };
    auto side_d = [](const RowR& r, int s, R val) { return s == 0 ? val - r.lb : r.ub - val; };
    auto side_eps = [](const RowR& r) { return r.hard ? R(1.0e-6) : R(1) / r.rho; };
    auto is_eq = [](const RowR& r) { return r.kind == EQF || r.kind == EQX; };
    long nsides = 0;
    for (int k = 0; k <= N; ++k)
        for (RowR& r : st[k].rows) {
            const R val = row_val(r, z[k].data());
            for (int s = 0; s < 2; ++s) {
                if (!r.on[s]) continue;
                r.t[s] = std::max(side_d(r, s, val), R(P.qp_thr0));
                r.lam[s] = R(P.qp_mu0) / r.t[s];
                ++nsides;
            }
        }
    // per-stage force / equality block
    struct FBlock {
        int ne = 0;
        std::vector<int> rows;             // indices of the EQF rows
        std::vector<ST> Dinv;              // nfc x nfc
        std::vector<ST> T;                 // Df D^-1, ne x nfc
        std::vector<ST> L;                 // Cholesky factor of S, ne x ne (lower)
        std::vector<ST> G;                 // L^-1 C, ne x nx
        std::vector<R> v;                  // e + y / rho
    };
    std::vector<FBlock> fb(N);
    std::vector<std::vector<F>> M(N + 1), grad(N + 1), dr;
    std::vector<std::vector<R>> rg(N + 1), dz(N + 1);
    R last_alpha = 0, last_step = std::numeric_limits<R>::infinity(), mu = 0, rd_max = 0;
    int iters = 0, failed = 0;
    bool converged = false;
    auto chol_st = [](std::vector<ST>& A_, int n) {   // in place, lower; returns false on a non-positive pivot
        for (int j = 0; j < n; ++j) {
            ST d = A_[j * n + j];
            for (int l = 0; l < j; ++l) d -= A_[j * n + l] * A_[j * n + l];
            if (!(d > ST(0))) return false;
            d = std::sqrt(d);
            A_[j * n + j] = d;
            for (int i = j + 1; i < n; ++i) {
                ST s_ = A_[i * n + j];
                for (int l = 0; l < j; ++l) s_ -= A_[i * n + l] * A_[j * n + l];
                A_[i * n + j] = s_ / d;
            }
        }
        return true;
    };
    for (int it = 0; it < P.qp_iter_max; ++it) {
        mu = 0;
        rd_max = 0;
        R pinf = 0;
        bool ok = true;
        for (int k = 0; k <= N; ++k) {
            StageR& s = st[k];
            const int nz = s.nz, xo = s.nu, jn = s.nu > 0 ? nq : 0;
            const int nrk = jn + nx;                       // reduced stage size (terminal: nx)
            auto ridx = [&](int i) { return i < jn ? i : i - (xo - jn); };   // stage index (jerk / state) -> reduced index
            M[k].assign(size_t(nrk) * nrk, F(0));
            for (int i = 0; i < nz; ++i) {
                if (i >= jn && i < xo) continue;
                for (int j = 0; j < nz; ++j) {
                    if (j >= jn && j < xo) continue;
                    M[k][ridx(i) * nrk + ridx(j)] = F(s.H[i * nz + j]);
                }
            }
            rg[k].assign(nz, R(0));
            for (int i = 0; i < nz; ++i) {
                R v = s.g[i];
                for (int j = 0; j < nz; ++j) v += s.H[i * nz + j] * z[k][j];
                rg[k][i] = v;
            }
            std::vector<ST> Dm;
            FBlock* B_ = k < N ? &fb[k] : nullptr;
            if (B_) {
                Dm.assign(size_t(nfc) * nfc, ST(0));
                for (int i = 0; i < nfc; ++i)
                    for (int j = 0; j < nfc; ++j) Dm[i * nfc + j] = ST(s.H[(nq + i) * nz + nq + j]);
                B_->rows.clear();
                B_->v.clear();
            }
            for (size_t ri = 0; ri < s.rows.size(); ++ri) {
                RowR& r = s.rows[ri];
                const R val = row_val(r, z[k].data());
                if (is_eq(r)) {
                    const R e = val - r.lb;
                    if (r.rho <= R(0)) continue;
                    if (r.hard) pinf = std::max(pinf, R(std::fabs(e)));
                    if (r.kind == EQF) {
                        B_->rows.push_back(int(ri));
                        B_->v.push_back(e + r.lambda / r.rho);
                    } else {   // state-only equality rows (terminal): penalty in the state block
                        for (int i = xo; i < nz; ++i) {
                            if (r.a[i] == R(0)) continue;
                            const F wi = F(r.rho) * F(r.a[i]);
                            for (int j = xo; j < nz; ++j) M[k][ridx(i) * nrk + ridx(j)] += wi * F(r.a[j]);
                            rg[k][i] += (r.rho * e + r.lambda) * r.a[i];
                        }
                    }
                    continue;
                }
                const R eps = side_eps(r);
                for (int sd = 0; sd < 2; ++sd) {
                    if (!r.on[sd]) continue;
                    const R rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                    rd_max = std::max(rd_max, R(std::fabs(rd)));
                    mu += r.t[sd] * r.lam[sd];
                    const R w = r.lam[sd] / (r.t[sd] + eps * r.lam[sd]);
                    if (r.idx >= 0) {
                        if (r.kind == FBOX) Dm[(r.idx - nq) * nfc + (r.idx - nq)] += ST(w);
                        else M[k][ridx(r.idx) * nrk + ridx(r.idx)] += F(w);
                        rg[k][r.idx] += -sgn[sd] * r.lam[sd];
                    } else {
                        if (r.kind == FRIC) {
                            for (int i = 0; i < nfc; ++i) {
                                if (r.a[nq + i] == R(0)) continue;
                                for (int j = 0; j < nfc; ++j) Dm[i * nfc + j] += ST(w) * ST(r.a[nq + i]) * ST(r.a[nq + j]);
                            }
                        } else {
                            for (int i = xo; i < nz; ++i) {
                                if (r.a[i] == R(0)) continue;
                                const F wi = F(w) * F(r.a[i]);
                                for (int j = xo; j < nz; ++j) M[k][ridx(i) * nrk + ridx(j)] += wi * F(r.a[j]);
                            }
                        }
                        for (int j = 0; j < nz; ++j) rg[k][j] += -sgn[sd] * r.lam[sd] * r.a[j];
                    }
                }
            }
            if (B_) B_->ne = int(B_->rows.size());
            if (B_ && nfc > 0) {
                const int ne = B_->ne;
                // D^-1 by Cholesky (block diagonal in the kernels)
                std::vector<ST> Lc = Dm;
                if (!chol_st(Lc, nfc)) { ok = false; break; }
                B_->Dinv.assign(size_t(nfc) * nfc, ST(0));
                for (int c = 0; c < nfc; ++c) {   // solve L L' x = e_c
                    std::vector<ST> y(nfc, ST(0));
                    for (int i = 0; i < nfc; ++i) {
                        ST s_ = (i == c) ? ST(1) : ST(0);
                        for (int l = 0; l < i; ++l) s_ -= Lc[i * nfc + l] * y[l];
                        y[i] = s_ / Lc[i * nfc + i];
                    }
                    for (int i = nfc - 1; i >= 0; --i) {
                        ST s_ = y[i];
                        for (int l = i + 1; l < nfc; ++l) s_ -= Lc[l * nfc + i] * B_->Dinv[l * nfc + c];
                        B_->Dinv[i * nfc + c] = s_ / Lc[i * nfc + i];
                    }
                }
                B_->T.assign(size_t(ne) * nfc, ST(0));
                for (int a = 0; a < ne; ++a) {
                    const RowR& r = s.rows[B_->rows[a]];
                    for (int j = 0; j < nfc; ++j) {
                        ST s_ = 0;
                        for (int l = 0; l < nfc; ++l) s_ += ST(r.a[nq + l]) * B_->Dinv[l * nfc + j];
                        B_->T[a * nfc + j] = s_;
                    }
                }
                B_->L.assign(size_t(ne) * ne, ST(0));
                for (int a = 0; a < ne; ++a)
                    for (int b2 = 0; b2 <= a; ++b2) {
                        const RowR& rb = s.rows[B_->rows[b2]];
                        ST s_ = (a == b2) ? ST(1) / ST(s.rows[B_->rows[a]].rho) : ST(0);
                        for (int j = 0; j < nfc; ++j) s_ += B_->T[a * nfc + j] * ST(rb.a[nq + j]);
                        B_->L[a * ne + b2] = s_;
                    }
                if (ne > 0 && !chol_st(B_->L, ne)) { ok = false; break; }
                B_->G.assign(size_t(ne) * nx, ST(0));
                for (int j = 0; j < nx; ++j)
                    for (int a = 0; a < ne; ++a) {
                        ST s_ = ST(s.rows[B_->rows[a]].a[xo + j]);
                        for (int l = 0; l < a; ++l) s_ -= B_->L[a * ne + l] * B_->G[l * nx + j];
                        B_->G[a * nx + j] = s_ / B_->L[a * ne + a];
                    }
                for (int i = 0; i < nx; ++i)
                    for (int j = 0; j < nx; ++j) {
                        F s_ = 0;
                        for (int a = 0; a < ne; ++a) s_ += F(B_->G[a * nx + i]) * F(B_->G[a * nx + j]);
                        M[k][(jn + i) * nrk + jn + j] += s_;
                    }
            }
        }
        mu = nsides > 0 ? mu / R(nsides) : R(0);
        if (ok && it > 0 && mu <= R(mu_factor * P.qp_mu_target) && rd_max <= R(P.qp_tol) && last_alpha >= R(0.5) &&
            (pinf <= R(P.qp_tol) || last_step <= R(P.qp_tol))) {
            converged = true;
            if (extra_left-- <= 0) break;
        }
        iters = it + 1;
        if (!ok || !ric.factor(M)) {
            failed = it + 1;
            break;
        }
        auto solve_with = [&](bool corrector, R target) {
            std::vector<std::vector<ST>> glam(N);
            std::vector<std::vector<R>> gfull(N + 1);
            for (int k = 0; k <= N; ++k) {
                StageR& s = st[k];
                const int nz = s.nz, xo = s.nu, jn = s.nu > 0 ? nq : 0;
                std::vector<R> gk = rg[k];
                for (RowR& r : s.rows) {
                    if (is_eq(r)) continue;
                    const R val = row_val(r, z[k].data());
                    const R eps = side_eps(r);
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!r.on[sd]) continue;
                        const R rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                        R rc = r.t[sd] * r.lam[sd] - target;
                        if (corrector) rc += r.dt[sd] * r.dl[sd];
                        const R w = sgn[sd] * (rc + r.lam[sd] * rd) / (r.t[sd] + eps * r.lam[sd]);
                        if (r.idx >= 0) gk[r.idx] += w;
                        else for (int j = 0; j < nz; ++j) gk[j] += w * r.a[j];
                    }
                }
                gfull[k] = gk;
                grad[k].assign(jn + nx, F(0));
                for (int i = 0; i < jn; ++i) grad[k][i] = F(gk[i]);
                std::vector<R> gx(gk.begin() + xo, gk.end());
                if (k < N && fb[k].ne > 0) {
                    const FBlock& B_ = fb[k];
                    const int ne = B_.ne;
                    std::vector<ST> rhs(ne);
                    for (int a = 0; a < ne; ++a) {
                        R s_ = B_.v[a];
                        for (int j = 0; j < nfc; ++j) s_ -= R(B_.T[a * nfc + j]) * gk[nq + j];
                        rhs[a] = ST(s_);
                    }
                    glam[k].assign(ne, ST(0));
                    for (int a = 0; a < ne; ++a) {
                        ST s_ = rhs[a];
                        for (int l = 0; l < a; ++l) s_ -= B_.L[a * ne + l] * glam[k][l];
                        glam[k][a] = s_ / B_.L[a * ne + a];
                    }
                    for (int j = 0; j < nx; ++j) {
                        R s_ = 0;
                        for (int a = 0; a < ne; ++a) s_ += R(B_.G[a * nx + j]) * R(glam[k][a]);
                        gx[j] += s_;
                    }
                }
                for (int i = 0; i < nx; ++i) grad[k][jn + i] = F(gx[i]);
            }
            ric.solve(grad, dr);
            for (int k = 0; k <= N; ++k) {
                StageR& s = st[k];
                const int xo = s.nu, jn = s.nu > 0 ? nq : 0;
                dz[k].assign(s.nz, R(0));
                for (int i = 0; i < jn; ++i) dz[k][i] = R(dr[k][i]);
                for (int i = 0; i < nx; ++i) dz[k][xo + i] = R(dr[k][jn + i]);
                if (k < N && nfc > 0) {
                    const FBlock& B_ = fb[k];
                    const int ne = B_.ne;
                    std::vector<ST> lamv(ne, ST(0));
                    if (ne > 0) {
                        std::vector<ST> t(ne);
                        for (int a = 0; a < ne; ++a) {
                            ST s_ = glam[k][a];
                            for (int j = 0; j < nx; ++j) s_ += B_.G[a * nx + j] * ST(dr[k][jn + j]);
                            t[a] = s_;
                        }
                        for (int a = ne - 1; a >= 0; --a) {
                            ST s_ = t[a];
                            for (int l = a + 1; l < ne; ++l) s_ -= B_.L[l * ne + a] * lamv[l];
                            lamv[a] = s_ / B_.L[a * ne + a];
                        }
                    }
                    std::vector<R> q(nfc);
                    for (int j = 0; j < nfc; ++j) {
                        R s_ = gfull[k][nq + j];
                        for (int a = 0; a < ne; ++a) s_ += s.rows[B_.rows[a]].a[nq + j] * R(lamv[a]);
                        q[j] = s_;
                    }
                    for (int i = 0; i < nfc; ++i) {
                        R s_ = 0;
                        for (int j = 0; j < nfc; ++j) s_ -= R(B_.Dinv[i * nfc + j]) * q[j];
                        dz[k][nq + i] = s_;
                    }
                }
            }
            for (int k = 0; k <= N; ++k)
                for (RowR& r : st[k].rows) {
                    if (is_eq(r)) continue;
                    const R val = row_val(r, z[k].data());
                    const R eps = side_eps(r);
                    R adz = 0;
                    if (r.idx >= 0) adz = dz[k][r.idx];
                    else for (size_t j = 0; j < r.a.size(); ++j) adz += r.a[j] * dz[k][j];
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!r.on[sd]) continue;
                        const R rd = side_d(r, sd, val) + eps * r.lam[sd] - r.t[sd];
                        R rc = r.t[sd] * r.lam[sd] - target;
                        if (corrector) rc += r.dt[sd] * r.dl[sd];
                        const R den = r.t[sd] + eps * r.lam[sd];
                        const R dl = -(rc + r.lam[sd] * rd) / den - (r.lam[sd] / den) * sgn[sd] * adz;
                        r.dl[sd] = dl;
                        r.dt[sd] = sgn[sd] * adz + eps * dl + rd;
                    }
                }
        };
        auto max_step = [&]() {
            R a = 1;
            for (int k = 0; k <= N; ++k)
                for (const RowR& r : st[k].rows) {
                    if (is_eq(r)) continue;
                    for (int sd = 0; sd < 2; ++sd) {
                        if (!r.on[sd]) continue;
                        if (r.dt[sd] < R(0)) a = std::min(a, -r.t[sd] / r.dt[sd]);
                        if (r.dl[sd] < R(0)) a = std::min(a, -r.lam[sd] / r.dl[sd]);
                    }
                }
            return a;
        };
        for (int k = 0; k <= N; ++k)
            for (RowR& r : st[k].rows) r.dt[0] = r.dt[1] = r.dl[0] = r.dl[1] = R(0);
        if (nsides > 0) {
            solve_with(false, R(0));
            const R a_aff = max_step();
            R mu_aff = 0;
            for (int k = 0; k <= N; ++k)
                for (const RowR& r : st[k].rows) {
                    if (is_eq(r)) continue;
                    for (int sd = 0; sd < 2; ++sd)
                        if (r.on[sd]) mu_aff += (r.t[sd] + a_aff * r.dt[sd]) * (r.lam[sd] + a_aff * r.dl[sd]);
                }
            mu_aff /= R(nsides);
            const R ratio = mu_aff / mu;
            solve_with(true, std::max(ratio * ratio * ratio * mu, R(P.qp_mu_target)));
        } else {
            solve_with(false, R(0));
        }
        const R alpha = nsides > 0 ? std::min(R(1), R(0.995) * max_step()) : R(1);
        last_step = 0;
        for (int k = 0; k <= N; ++k) {
            for (int i = 0; i < st[k].nz; ++i) {
                const R step = alpha * dz[k][i];
                z[k][i] += step;
                last_step = std::max(last_step, R(std::fabs(step)));
            }
            for (RowR& r : st[k].rows) {
                if (is_eq(r)) {
                    if (r.hard && r.rho > R(0)) r.lambda += r.rho * (row_val(r, z[k].data()) - r.lb);
                    continue;
                }
                for (int sd = 0; sd < 2; ++sd) {
                    if (!r.on[sd]) continue;
                    r.t[sd] += alpha * r.dt[sd];
                    r.lam[sd] += alpha * r.dl[sd];
                }
            }
        }
        last_alpha = alpha;
        if (!(last_step < std::numeric_limits<R>::infinity())) {
            failed = it + 1;
            break;
        }
    }
    zout.assign(N + 1, {});
    for (int k = 0; k <= N; ++k) zout[k].assign(z[k].begin(), z[k].end());
    info[0] = iters; info[1] = converged; info[2] = failed; info[3] = double(mu); info[4] = double(rd_max);
}
