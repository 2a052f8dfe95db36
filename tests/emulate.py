"""numpy emulation of the device pipeline driven by the HOST-BUILT tables of
libmsfec_b200.so (msfec_debug_table).  Test infrastructure: lets the CPU-only test
suite check the topology / assembly / operator tables (what the CUDA kernels consume)
against the oracle without a GPU.  The arithmetic mirrors csrc/engine.cu kernel by
kernel; the Krylov solve is replaced by a sparse direct solve."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import msfec_oracle as mo


def dims_of(bb):
    import importlib
    d = bb.table("dims")
    from conftest import msfec_mod
    return dict(zip(msfec_mod().DIMS, d.tolist()))


def coef_channels(prob: "mo.Problem", corners, cell_gid, D):
    x0 = np.asarray(corners)[0]
    H = corners[7][0] - x0[0]
    A, Ainv, B, pts = mo.coefficient_fields(prob, x0, H, cell_gid)
    Tn = Ainv if D["tensor_inverse"] else A
    s = 1.0 / B if D["scalar_inverse"] else B
    nC = Tn.shape[0]
    coef = np.zeros((nC, 8, 7))
    c = 0
    for a in range(3):
        for b in range(a, 3):
            coef[:, :, c] = Tn[:, :, a, b]; c += 1
    coef[:, :, 6] = s
    exprs = [mo.Expr(e, prob.rhs_constants) for e in prob.rhs_expr.split(";")]
    fr = np.stack([e(pts) for e in exprs], -1)          # [nC, 8, ncomp]
    return coef.reshape(nC, 56), fr.reshape(nC, -1), H


def assemble_slots(bb, prefix, coef, scale):
    cp = bb.table(prefix + ".contrib_ptr"); cc = bb.table(prefix + ".contrib_cell"); cpair = bb.table(prefix + ".contrib_pair")
    pp = bb.table(prefix + ".pair_ptr"); pi = bb.table(prefix + ".pair_idx"); pw = bb.table(prefix + ".pair_w")
    n_slots = len(cp) - 1
    out = np.zeros(n_slots)
    # per-contribution value
    npairs = len(pp) - 1
    # value of pair p on every fine cell: sum_e w_e coef[T, idx_e]
    pv = np.zeros((coef.shape[0], npairs))
    for p in range(npairs):
        sl = slice(pp[p], pp[p + 1])
        if pp[p + 1] > pp[p]:
            pv[:, p] = coef[:, pi[sl]] @ pw[sl]
    vals = pv[cc, cpair]
    slot_of = np.repeat(np.arange(n_slots), np.diff(cp))
    np.add.at(out, slot_of, vals)
    return out * scale


def ref_operator(bb, name, vals, kscale, shape):
    cptr = bb.table(name + ".cptr"); ccol = bb.table(name + ".ccol"); cref = bb.table(name + ".cref")
    sptr = bb.table(name + ".sptr"); scol = bb.table(name + ".scol"); sval = bb.table(name + ".sval")
    n_rows = len(cptr) - 1
    rows_c = np.repeat(np.arange(n_rows), np.diff(cptr))
    v_c = np.where(cref & 1, -1.0, 1.0) * vals[cref >> 1] if len(cref) else np.zeros(0)
    rows_s = np.repeat(np.arange(n_rows), np.diff(sptr))
    A = sp.coo_matrix((np.concatenate([v_c, sval * kscale]),
                       (np.concatenate([rows_c, rows_s]), np.concatenate([ccol, scol]))), shape=shape)
    return A.tocsr()


def emulate_cell(bb, prob, corners, cell_gid=0):
    """Returns M, r, Z (full basis store [NF, kg]) computed from the library's tables."""
    D = dims_of(bb)
    coef, fr, H = coef_channels(prob, corners, cell_gid, D)
    h = H / D["n"]
    v00 = assemble_slots(bb, "asm00", coef, h ** D["asm00_h_exponent"] / 8.0)
    v11 = assemble_slots(bb, "asm11", coef, h ** D["asm11_h_exponent"] / 8.0) if D["n_slots1"] else np.zeros(0)
    vals = np.concatenate([v00, v11])
    grhs = assemble_slots(bb, "asm_rhs", fr, h ** D["asm_rhs_h_exponent"] / 8.0)
    kscale = h ** D["k_h_exponent"]; f1scale = H ** D["f1_H_exponent"]
    NI, NB, NF, k, kg, k0 = D["NI"], D["NB"], D["NF"], D["k_solve"], D["k_gram"], D["k0"]
    Sys = ref_operator(bb, "sys", vals, kscale, (NI, NI))
    Lift = ref_operator(bb, "lift", vals, kscale, (NI, NB))
    Full = ref_operator(bb, "full", vals, kscale, (NF, NF))
    G = bb.table("G").reshape(k, NB); F1 = bb.table("F1").reshape(k, NI)
    b = F1.T * f1scale + Lift @ G.T
    if D["pairing"] == 3:   # RT_DQ: singular (constant u); pin last unknown for the direct emulation
        keep = np.arange(NI - 1)
        x = np.zeros((NI, k))
        x[keep] = spla.splu(Sys[keep][:, keep].tocsc()).solve(b[keep])
    else:
        x = spla.splu(Sys.tocsc()).solve(b)
    NI0, N0, NI1, N1 = D["NI0"], D["N0"], D["NI1"], D["N1"]
    NB0 = N0 - NI0
    Z = np.zeros((NF, kg))
    for j in range(kg):
        if j < k0:
            Z[:NI0, j] = x[:NI0, j]; Z[NI0:N0, j] = G[j, :NB0]
        elif D["pairing"] == 3:
            Z[N0:, j] = 1.0
        else:
            Z[N0:N0 + NI1, j] = x[NI0:, j]; Z[N0 + NI1:, j] = G[j, NB0:]
    M = Z.T @ (Full @ Z)
    off = N0 if D["rhs_block"] else 0
    r = Z[off:off + len(grhs)].T @ grhs
    return M, r, Z, dict(vals=vals, grhs=grhs, Sys=Sys, b=b, x=x)


def emulate_direct(bb, vals, kscale, b):
    """numpy model of the batched direct solver driven by the library's DirectPlan tables: fills the per-cell
    'band' exactly like k_direct_fill_* (cell_dest/cell_ref, shared_dest/shared_val, const_dest/const_val,
    rhs_dest), decodes it into the padded symmetric matrix through the front tables (col_off, ld, chunk_*),
    runs a dense LDL^T WITHOUT pivoting in the padded elimination order and returns (x in interior numbering,
    pivots d, padded->interior map)."""
    bs = bb.table("direct.bs"); off = bb.table("direct.slab_off"); ld = bb.table("direct.ld")
    col_off = bb.table("direct.col_off"); front_rows = bb.table("direct.front_rows")
    ch_off = bb.table("direct.chunk_off"); ch_blk = bb.table("direct.chunk_blk"); ch_loc = bb.table("direct.chunk_local")
    inv = bb.table("direct.inv_perm"); NP = len(inv); k = b.shape[1]
    band = np.zeros(int(col_off[-1] + ld[-1] * bs[-1]))
    cd = bb.table("direct.cell_dest"); cr = bb.table("direct.cell_ref")
    band[cd] = np.where(cr & 1, -1.0, 1.0) * vals[cr >> 1]
    sd = bb.table("direct.shared_dest")
    if len(sd):
        band[sd] = bb.table("direct.shared_val") * kscale
    kd = bb.table("direct.const_dest")
    if len(kd):
        band[kd] = bb.table("direct.const_val")
    rd = bb.table("direct.rhs_dest")
    for r in range(len(rd)):
        if rd[r] >= 0:
            band[rd[r]:rd[r] + k] = b[r]
    # decode band -> padded matrix A (lower) and rhs F
    A = np.zeros((NP, NP)); F = np.zeros((NP, k))
    for s in range(len(bs)):
        P = band[col_off[s]: col_off[s] + ld[s] * bs[s]].reshape(bs[s], ld[s])      # [col][row] (column-major)
        for t in range(ch_off[s], ch_off[s + 1]):
            v0 = (t - ch_off[s]) * 32
            if ch_blk[t] < 0:
                F[off[s]: off[s] + bs[s], :] = P[:, v0: v0 + k]
            else:
                r0 = off[ch_blk[t]] + ch_loc[t]
                A[r0: r0 + 32, off[s]: off[s] + bs[s]] = P[:, v0: v0 + 32].T
    A = np.tril(A) + np.tril(A, -1).T
    L = np.zeros_like(A); d = np.zeros(NP); W = A.copy()
    for j in range(NP):
        d[j] = W[j, j]; L[j:, j] = W[j:, j] / d[j]
        W[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j]) * d[j]
    xp = np.linalg.solve(L.T, np.linalg.solve(L, F) / d[:, None])
    x = np.zeros_like(b)
    real = inv >= 0
    x[inv[real]] = xp[real]
    return x, d, inv


MF_FIELDS = ["s", "s8", "u", "u8", "m", "ldx", "level", "parent", "own_base", "idx_off", "row_off", "pe_lo", "pe_hi",
             "ps_lo", "ps_hi", "pc_lo", "pc_hi", "ch_lo", "ch_hi", "l_off", "c_off", "n_rows_real", "ldc_max"]


def mf_tables(bb):
    info = bb.table("mf.info")
    nfld = int(info[9])
    assert nfld == len(MF_FIELDS)
    fr = bb.table("mf.fronts").reshape(-1, nfld)
    T = dict(info=dict(feasible=bool(info[0]), kr=int(info[1]), NP=int(info[2]), n_levels=int(info[3]),
                       l_doubles=int(info[4]), c_doubles=int(info[5]), flops=info[6], bytes=info[7], pinned_row=int(info[8])),
             fronts=[dict(zip(MF_FIELDS, row.tolist())) for row in fr],
             children=bb.table("mf.children").reshape(-1, 4))
    for name in ("front_idx", "own_rows", "cmap", "pinv", "pe_dest", "pe_ref", "ps_dest", "ps_val", "pc_dest", "pc_val",
                 "perm", "inv_perm", "level_off", "level_fronts", "smem_fwd", "smem_bwd", "smem_fwd_st", "rt_max"):
        T[name] = bb.table("mf." + name)
    return T


def emulate_multifrontal(bb, vals, kscale, b):
    """numpy model of the multifrontal kernels (csrc/mf.cuh) driven by the library's MfPlan tables: level by level,
    every front assembles its panel [m x s8] from the per-cell slot values, the shared entries, the rhs and the first
    n_own columns of its children's contribution blocks; factors it (LDL^T without pivoting, X = L D kept for the
    rows below); stores the factor panel; forms its contribution block C (column-major (u8+kr) x u8) by gathering the
    children through `pinv` and subtracting X L21^T.  Backward pass top-down through front_idx.  Returns
    (x in interior numbering, pivots d in padded numbering, padded->interior map)."""
    T = mf_tables(bb)
    assert T["info"]["feasible"]
    kr, NP = T["info"]["kr"], T["info"]["NP"]
    k = b.shape[1]
    Lst = np.zeros(T["info"]["l_doubles"]); Cst = np.full(T["info"]["c_doubles"], np.nan)
    dall = np.ones(NP)
    fronts, ch = T["fronts"], T["children"]
    order_fwd = [f for l in range(T["info"]["n_levels"]) for f in T["level_fronts"][T["level_off"][l]:T["level_off"][l + 1]]]
    for f in order_fwd:
        F = fronts[f]
        s8, u8, m, ldx = F["s8"], F["u8"], F["m"], F["ldx"]
        P = np.zeros(m * ldx)
        sl = slice(F["pe_lo"], F["pe_hi"]); ref = T["pe_ref"][sl]
        P[T["pe_dest"][sl]] = np.where(ref & 1, -1.0, 1.0) * vals[ref >> 1]
        sl = slice(F["ps_lo"], F["ps_hi"]); P[T["ps_dest"][sl]] = T["ps_val"][sl] * kscale
        sl = slice(F["pc_lo"], F["pc_hi"]); P[T["pc_dest"][sl]] = T["pc_val"][sl]
        P = P.reshape(m, ldx)
        rows = T["own_rows"][F["row_off"]:F["row_off"] + s8]
        for c in range(s8):
            if rows[c] >= 0:
                P[s8 + u8:s8 + u8 + k, c] = b[rows[c]]
        kids = ch[F["ch_lo"]:F["ch_hi"]]
        for (cf, n_own, cmap_off, pinv_off) in kids:
            Cf = fronts[cf]
            ldc = Cf["u8"] + kr
            Cc = Cst[Cf["c_off"]:Cf["c_off"] + ldc * Cf["u8"]].reshape(Cf["u8"], ldc)      # [col][row]
            cmap = T["cmap"][cmap_off:cmap_off + ldc]
            for j in range(n_own):
                assert 0 <= cmap[j] < s8
                for i in range(j, ldc):
                    if cmap[i] >= 0:
                        P[cmap[i], cmap[j]] += Cc[j, i]
        # factor the panel: right-looking LDL^T; afterwards the own block holds the unit-lower L (strictly lower part),
        # the rows below hold X = L D
        X = P[:, :s8].copy()
        d = np.zeros(s8)
        for j in range(s8):
            d[j] = X[j, j]
            lcol_own = X[j + 1:s8, j] / d[j]                       # L(i, j) for own rows i > j
            X[j + 1:, j + 1:s8] -= np.outer(X[j + 1:, j], lcol_own)
            X[j + 1:s8, j] = lcol_own
        Lfull = X.copy()
        Lfull[s8:, :] = X[s8:, :] / d[None, :]
        # factor record = shared-memory image of k_mf_forward: panel [m][ldx] (rows below a pivot tile hold X = L D, the
        # pivot tiles themselves are not read again), 1/d, d, unit-lower factors of the 8 x 8 pivot tiles
        rec = np.zeros(m * ldx + 10 * s8)
        Pn = rec[:m * ldx].reshape(m, ldx)
        Pn[s8:, :s8] = X[s8:, :]
        Ld = np.zeros((s8, 8))
        for r in range(s8):
            for c in range(r):
                if r // 8 == c // 8:
                    Ld[r, c % 8] = X[r, c]                       # unit-lower L inside the pivot tile
                else:
                    Pn[r, c] = X[r, c] * d[c]                    # X = L D below the pivot tile
            Pn[r, (r // 8) * 8:(r // 8) * 8 + 8] = np.nan        # pivot tile: garbage in the kernel, never read
        rec[m * ldx:m * ldx + s8] = 1.0 / d
        rec[m * ldx + s8:m * ldx + 2 * s8] = d
        # the record keeps the INVERSES of the unit-lower pivot tiles (the rows below are solved with them on the tensor cores,
        # and so is the back-substitution)
        Li = np.zeros((s8, 8))
        for q in range(s8 // 8):
            Li[q * 8:q * 8 + 8] = np.linalg.inv(np.eye(8) + Ld[q * 8:q * 8 + 8])
        rec[m * ldx + 2 * s8:] = Li.reshape(-1)
        Lst[F["l_off"]:F["l_off"] + len(rec)] = rec
        dall[F["own_base"]:F["own_base"] + s8] = d
        # contribution block
        if u8:
            ldc = u8 + kr
            Cn = np.zeros((u8, ldc))
            for (cf, n_own, cmap_off, pinv_off) in kids:
                Cf = fronts[cf]
                ldcc = Cf["u8"] + kr
                Cc = Cst[Cf["c_off"]:Cf["c_off"] + ldcc * Cf["u8"]].reshape(Cf["u8"], ldcc)
                pinv = T["pinv"][pinv_off:pinv_off + m]
                for j in range(u8):
                    jc = pinv[s8 + j]
                    if jc < 0:
                        continue
                    for i in range(j, ldc):
                        ic = pinv[s8 + i]
                        if ic >= 0:
                            assert ic >= jc
                            Cn[j, i] += Cc[jc, ic]
            Xb = X[s8:, :]                       # (u8 + kr) x s8
            Lb = Lfull[s8:s8 + u8, :]            # u8 x s8
            Cn -= (Xb @ Lb.T).T
            for j in range(u8):
                Cn[j, :j] = np.nan               # entries above the diagonal are never written by the kernel
            Cst[F["c_off"]:F["c_off"] + ldc * u8] = Cn.reshape(-1)
    # backward
    xp = np.zeros((NP, kr))
    for f in reversed(order_fwd):
        F = fronts[f]
        s8, u8, m = F["s8"], F["u8"], F["m"]
        ldx = F["ldx"]
        rec = Lst[F["l_off"]:F["l_off"] + m * ldx + 10 * s8]
        Pn = rec[:m * ldx].reshape(m, ldx)
        dinv = rec[m * ldx:m * ldx + s8]
        Li = rec[m * ldx + 2 * s8:].reshape(s8, 8)
        idx = T["front_idx"][F["idx_off"]:F["idx_off"] + s8 + u8]
        xu = np.zeros((u8, kr))
        for r in range(u8):
            if idx[s8 + r] >= 0:
                xu[r] = xp[idx[s8 + r]]
        t = (Pn[s8 + u8:, :s8].T - Pn[s8:s8 + u8, :s8].T @ xu) * dinv[:, None]
        # L11^T x = t tile by tile, last tile first: x_p = Linv_p^T t_p, then t_c -= (1 / d_c) X(p rows, c)^T x_p for c before p
        xo = t.copy()
        for q in range(s8 // 8 - 1, -1, -1):
            c0 = q * 8
            xo[c0:c0 + 8] = Li[c0:c0 + 8].T @ xo[c0:c0 + 8]
            xo[:c0] -= (Pn[c0:c0 + 8, :c0].T @ xo[c0:c0 + 8]) * dinv[:c0, None]
        xp[F["own_base"]:F["own_base"] + s8] = xo
    inv = T["inv_perm"]
    x = np.zeros_like(b)
    real = inv >= 0
    x[inv[real]] = xp[real][:, :k]
    return x, dall, inv
