"""CPU oracle for the MsFEC per-coarse-cell multiscale basis build.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (msfec_b200/, csrc/) may
import, call, link or execute this file; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs do.

This is a deal.II-free *restatement* (numpy + scipy SuperLU) of what the
reference's four ``*Basis::run()`` pipelines compute when
``use direct solver basis = true`` (the exact discrete solution):

  reference (all paths relative to /root/reference)        here
  -------------------------------------------------------  --------------------
  source/Ned_RT/ned_rt_basis.cc:362-576  assemble_system    assemble_blocks()
  source/Q/q_basis.cc:175-257                                  "
  source/Q_Ned/q_ned_basis.cc:360-567                          "
  source/RT_DQ/rt_dq_basis.cc:391-579                          "
  ned_rt_basis.cc:225-359  setup_basis_dofs_{curl,div}      boundary_data()
  ned_rt_basis.cc:1338-1371 condense + :579-634 solve_direct solve_basis()
  ned_rt_basis.cc:850-948 assemble_global_element_matrix    element_matrix()
  rt_dq_basis.cc:1039-1043 set_u_to_std, :822-913           element_matrix()
  source/equation_data/eqn_coeff_A.cc:142-242               CoefficientA
  source/equation_data/eqn_coeff_B.cc:72-114                CoefficientB
  source/equation_data/eqn_rhs.cc:58-107                    parsed rhs (Expr)
  include/functions/basis_{q1,q1_grad,nedelec,nedelec_curl,
      raviart_thomas}.tpp                                   coarse_* closed forms
  source/*/ *_parameters.cc + example_parameters/*.prm      parse_prm()

PARITY PINNING.  deal.II / Trilinos / UMFPACK are not vendored in the reference
and not installed here, so the reference cannot be run.  The only golden data the
reference's own tests hold at this boundary is
test/test_fe_projection_nedelec_mpi.mpirun=1.output (Nedelec DoF = tangential
component x edge length); tests/test_oracle.py checks the oracle's Nedelec
space against a digest of that file (tests/golden/).  Coarse element matrices
themselves are "parity unpinned" by the reference; they are pinned here by the
analytic invariants listed in SURVEY.md section 8(c).

Conventions (deal.II, restated in SURVEY.md App. A): hex vertex v = i+2j+4k,
lines 0..11 / faces 0..5 in GeometryInfo<3> order, Nedelec DoF = int_e u.t with
t in +coordinate direction, RT DoF = int_F u.n with n in +coordinate direction,
QGauss<3>(2) quadrature at the physical points.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

PAIRINGS = ("Q", "Q_NED", "NED_RT", "RT_DQ")

# --------------------------------------------------------------------------
# .prm reader (deal.II ParameterHandler text format, subset used by the
# reference: subsection/end, set key = value, # comments, '\' continuation)
# --------------------------------------------------------------------------


def parse_prm_text(text: str) -> dict:
    out: dict = {}
    stack = [out]
    pending = ""
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.endswith("\\"):
            pending += line[:-1] + " "
            continue
        line = (pending + line).strip()
        pending = ""
        low = line.lower()
        if low.startswith("subsection "):
            name = " ".join(line[len("subsection "):].split())
            stack.append(stack[-1].setdefault(name, {}))
        elif low == "end":
            if len(stack) == 1:
                raise ValueError("unbalanced 'end' in prm")
            stack.pop()
        elif low.startswith("set "):
            key, _, val = line[4:].partition("=")
            stack[-1][" ".join(key.split())] = val.strip()
        else:
            raise ValueError(f"cannot parse prm line: {raw!r}")
    if len(stack) != 1:
        raise ValueError("unterminated subsection in prm")
    return out


def parse_prm(path: str) -> dict:
    with open(path) as f:
        return parse_prm_text(f.read())


def _as_bool(s: str) -> bool:
    s = s.strip().lower()
    if s in ("true", "yes", "on"):
        return True
    if s in ("false", "no", "off"):
        return False
    raise ValueError(f"not a bool: {s!r}")


# --------------------------------------------------------------------------
# Expression evaluation (muParser subset as used through deal.II FunctionParser)
# --------------------------------------------------------------------------

_TOKEN = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)|([A-Za-z_][A-Za-z_0-9]*)|(.))")

_FUNCS1 = {
    "sin": np.sin, "cos": np.cos, "tan": np.tan, "asin": np.arcsin,
    "acos": np.arccos, "atan": np.arctan, "sinh": np.sinh, "cosh": np.cosh,
    "tanh": np.tanh, "exp": np.exp, "log": np.log, "ln": np.log,
    "log10": np.log10, "log2": np.log2, "sqrt": np.sqrt, "abs": np.abs,
    "sign": np.sign, "floor": np.floor, "ceil": np.ceil,
}
_FUNCS2 = {"pow": np.power, "min": np.minimum, "max": np.maximum}


class Expr:
    """Recursive-descent compiler of one scalar expression to an RPN program.

    Grammar (lowest to highest precedence): + - | * / | unary +- | ^ (right
    assoc., binds tighter than unary minus, as in muParser) | atoms.
    """

    def __init__(self, text: str, constants: dict | None = None,
                 variables=("x", "y", "z")):
        self.text = text
        self.constants = {"pi": math.pi, "Pi": math.pi, "_pi": math.pi,
                          "_e": math.e}
        if constants:
            self.constants.update(constants)
        self.variables = tuple(variables)
        self.toks = [m.groups() for m in _TOKEN.finditer(text) if any(m.groups())]
        self.pos = 0
        self.prog: list = []
        self._sum()
        if self.pos != len(self.toks):
            raise ValueError(f"trailing tokens in expression {text!r}")

    # -- parser -----------------------------------------------------------
    def _peek(self):
        return self.toks[self.pos] if self.pos < len(self.toks) else (None, None, None)

    def _sym(self, ch):
        t = self._peek()
        if t[2] == ch:
            self.pos += 1
            return True
        return False

    def _sum(self):
        self._prod()
        while True:
            if self._sym("+"):
                self._prod(); self.prog.append(("op", "+"))
            elif self._sym("-"):
                self._prod(); self.prog.append(("op", "-"))
            else:
                return

    def _prod(self):
        self._unary()
        while True:
            if self._sym("*"):
                self._unary(); self.prog.append(("op", "*"))
            elif self._sym("/"):
                self._unary(); self.prog.append(("op", "/"))
            else:
                return

    def _unary(self):
        if self._sym("-"):
            self._unary(); self.prog.append(("op", "neg"))
        elif self._sym("+"):
            self._unary()
        else:
            self._power()

    def _power(self):
        self._atom()
        if self._sym("^"):
            self._unary_pow(); self.prog.append(("op", "^"))

    def _unary_pow(self):
        if self._sym("-"):
            self._unary_pow(); self.prog.append(("op", "neg"))
        elif self._sym("+"):
            self._unary_pow()
        else:
            self._power()

    def _atom(self):
        num, name, sym = self._peek()
        if num is not None:
            self.pos += 1
            self.prog.append(("const", float(num)))
        elif name is not None:
            self.pos += 1
            if self._sym("("):
                nargs = 1
                self._sum()
                while self._sym(","):
                    self._sum(); nargs += 1
                if not self._sym(")"):
                    raise ValueError("missing ')' in " + self.text)
                if name in _FUNCS1 and nargs == 1:
                    self.prog.append(("f1", name))
                elif name in _FUNCS2 and nargs == 2:
                    self.prog.append(("f2", name))
                else:
                    raise ValueError(f"unknown function {name}/{nargs}")
            elif name in self.variables:
                self.prog.append(("var", self.variables.index(name)))
            elif name in self.constants:
                self.prog.append(("const", float(self.constants[name])))
            else:
                raise ValueError(f"unknown identifier {name!r} in {self.text!r}")
        elif sym == "(":
            self.pos += 1
            self._sum()
            if not self._sym(")"):
                raise ValueError("missing ')' in " + self.text)
        else:
            raise ValueError(f"unexpected token {sym!r} in {self.text!r}")

    # -- evaluator --------------------------------------------------------
    def __call__(self, pts: np.ndarray) -> np.ndarray:
        """pts[..., 3] -> values[...]"""
        st = []
        for kind, arg in self.prog:
            if kind == "const":
                st.append(np.full(pts.shape[:-1], arg))
            elif kind == "var":
                st.append(pts[..., arg].astype(float))
            elif kind == "f1":
                st.append(_FUNCS1[arg](st.pop()))
            elif kind == "f2":
                b = st.pop(); a = st.pop(); st.append(_FUNCS2[arg](a, b))
            elif arg == "neg":
                st.append(-st.pop())
            else:
                b = st.pop(); a = st.pop()
                st.append(a + b if arg == "+" else a - b if arg == "-" else
                          a * b if arg == "*" else a / b if arg == "/" else
                          np.power(a, b))
        assert len(st) == 1
        return st[0]


def parse_constants(s: str) -> dict:
    out = {}
    for item in s.split(","):
        item = item.strip()
        if item:
            k, _, v = item.partition("=")
            out[k.strip()] = float(v)
    return out


# --------------------------------------------------------------------------
# Problem description
# --------------------------------------------------------------------------

EULER = (math.pi / 3, math.pi / 6, math.pi / 4)   # eqn_coeff_A.cc:8-10


def rotation_matrix(rotate: bool) -> np.ndarray:
    """eqn_coeff_A.cc:25-56"""
    if not rotate:
        return np.eye(3)
    a, b, g = EULER
    ca, sa, cb, sb, cg, sg = math.cos(a), math.sin(a), math.cos(b), math.sin(b), math.cos(g), math.sin(g)
    return np.array([
        [ca * cg - sa * cb * sg, -ca * sg - sa * cb * cg, sa * sb],
        [sa * cg + ca * cb * sg, -sa * sg + ca * cb * cg, -ca * sb],
        [sb * sg, sb * cg, cb]])


@dataclass
class Problem:
    """Everything a basis build needs besides the coarse-cell corners."""
    pairing: str = "NED_RT"
    n_refine_local: int = 3
    n_refine_global: int = 2
    # Diffusion A (eqn_coeff_A.cc): diag(scale_d (1 - alpha_d sin(2 pi k_d x_d))), rotated
    a_freq: tuple = (0, 0, 0)
    a_scale: tuple = (1.0, 1.0, 1.0)
    a_alpha: tuple = (1.0, 1.0, 1.0)
    a_rotate: bool = True
    # Diffusion B (eqn_coeff_B.cc): parsed expression with constants pi, frequency, scale, alpha
    b_freq: int = 0
    b_scale: float = 1.0
    b_alpha: float = 1.0
    b_expr: str = "0"
    # Right-hand side (eqn_rhs.cc:58-107): n_components expressions separated by ';'
    rhs_expr: str = "0"
    rhs_constants: dict = field(default_factory=dict)
    # Harness-defined rough random field; 0 = off
    random_field_seed: int = 0
    random_field_sigma: float = math.log(10.0) / 2
    use_direct_solver_basis: bool = False
    verbose: bool = False

    @property
    def n(self) -> int:
        return 1 << self.n_refine_local

    @staticmethod
    def from_prm(path: str, pairing: str) -> "Problem":
        prm = parse_prm(path)
        ms = prm.get("Multiscale method parameters", {})
        eq = prm.get("Equation parameters", {})
        A = eq.get("Diffusion A", {})
        B = eq.get("Diffusion B", {})
        R = eq.get("Right-hand side", {})
        ctrl = ms.get("Control flow", {})
        return Problem(
            pairing=pairing,
            n_refine_local=int(ms.get("Mesh", {}).get("local refinements", 1)),
            n_refine_global=int(ms.get("Mesh", {}).get("global refinements", 2)),
            a_freq=tuple(int(A.get(f"frequency {d}", 0)) for d in "xyz"),
            a_scale=tuple(float(A.get(f"scale {d}", 1)) for d in "xyz"),
            a_alpha=tuple(float(A.get(f"alpha {d}", 1)) for d in "xyz"),
            a_rotate=_as_bool(A.get("rotate", "true")),
            b_freq=int(B.get("frequency", 0)),
            b_scale=float(B.get("scale", 1)),
            b_alpha=float(B.get("alpha", 1)),
            b_expr=B.get("Function expression", "0"),
            rhs_expr=R.get("Function expression", "0"),
            rhs_constants=parse_constants(R.get("Function constants", "")),
            use_direct_solver_basis=_as_bool(ctrl.get("use direct solver basis", "false")),
            verbose=_as_bool(ctrl.get("verbose basis", "false")),
        )


# --------------------------------------------------------------------------
# Counter-based RNG for the synthetic rough random field (harness-defined;
# the same integer hash is implemented in csrc/ and must stay bit-identical).
# --------------------------------------------------------------------------

_M64 = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def random_field_normals(seed: int, gidx: np.ndarray) -> np.ndarray:
    """Four N(0,1) variates per global fine-cell index (Box-Muller on splitmix64).

    gidx: uint64 array [...]; returns float64 [..., 4].  Partition independent:
    depends only on (seed, global fine cell index).
    """
    with np.errstate(over="ignore"):
        base = (gidx.astype(np.uint64) * np.uint64(4)) ^ (np.uint64(seed) * np.uint64(0xD1342543DE82EF95))
        u = np.stack([_splitmix64(base + np.uint64(c)) for c in range(4)], axis=-1)
    # 53-bit uniforms in (0,1]
    f = ((u >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / 9007199254740992.0)
    r0 = np.sqrt(-2.0 * np.log(f[..., 0])); r1 = np.sqrt(-2.0 * np.log(f[..., 2]))
    t0 = 2.0 * math.pi * f[..., 1]; t1 = 2.0 * math.pi * f[..., 3]
    return np.stack([r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1)


# --------------------------------------------------------------------------
# Fine grid topology: n^3 cells, shared by every coarse cell
# --------------------------------------------------------------------------


class FineGrid:
    def __init__(self, n: int):
        self.n = n
        n1 = n + 1
        self.nV = n1 ** 3
        self.nEx = n * n1 * n1
        self.nE = 3 * self.nEx
        self.nFx = n1 * n * n
        self.nF = 3 * self.nFx
        self.nC = n ** 3
        k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        i = i.ravel(); j = j.ravel(); k = k.ravel()          # cell (i,j,k), x fastest
        self.cell_ijk = np.stack([i, j, k], 1)
        V = lambda a, b, c: a + n1 * (b + n1 * c)
        EX = lambda a, b, c: a + n * (b + n1 * c)
        EY = lambda a, b, c: self.nEx + a + n1 * (b + n * c)
        EZ = lambda a, b, c: 2 * self.nEx + a + n1 * (b + n1 * c)
        FX = lambda a, b, c: a + n1 * (b + n * c)
        FY = lambda a, b, c: self.nFx + a + n * (b + n1 * c)
        FZ = lambda a, b, c: 2 * self.nFx + a + n * (b + n * c)
        self.cV = np.stack([V(i + (v & 1), j + ((v >> 1) & 1), k + (v >> 2)) for v in range(8)], 1)
        self.cE = np.stack([
            EY(i, j, k), EY(i + 1, j, k), EX(i, j, k), EX(i, j + 1, k),
            EY(i, j, k + 1), EY(i + 1, j, k + 1), EX(i, j, k + 1), EX(i, j + 1, k + 1),
            EZ(i, j, k), EZ(i + 1, j, k), EZ(i, j + 1, k), EZ(i + 1, j + 1, k)], 1)
        self.cF = np.stack([FX(i, j, k), FX(i + 1, j, k), FY(i, j, k), FY(i, j + 1, k),
                            FZ(i, j, k), FZ(i, j, k + 1)], 1)
        self.cC = np.arange(self.nC)[:, None]
        # entity positions in units of h (midpoints for edges, centres for faces)
        g = np.arange(n1)
        kk, jj, ii = np.meshgrid(g, g, g, indexing="ij")
        self.v_pos = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], 1).astype(float)
        self.v_bnd = ((self.v_pos == 0) | (self.v_pos == n)).any(1)

        def grid(nx, ny, nz, off):
            c, b, a = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
            return np.stack([a.ravel() + off[0], b.ravel() + off[1], c.ravel() + off[2]], 1).astype(float)

        ex = grid(n, n1, n1, (0.5, 0, 0)); ey = grid(n1, n, n1, (0, 0.5, 0)); ez = grid(n1, n1, n, (0, 0, 0.5))
        self.e_pos = np.concatenate([ex, ey, ez])
        self.e_dir = np.repeat(np.arange(3), self.nEx)
        fx = grid(n1, n, n, (0, 0.5, 0.5)); fy = grid(n, n1, n, (0.5, 0, 0.5)); fz = grid(n, n, n1, (0.5, 0.5, 0))
        self.f_pos = np.concatenate([fx, fy, fz])
        self.f_dir = np.repeat(np.arange(3), self.nFx)
        tr = np.ones((self.nE, 3), bool); tr[np.arange(self.nE), self.e_dir] = False
        self.e_bnd = (((self.e_pos == 0) | (self.e_pos == n)) & tr).any(1)
        nm = self.f_pos[np.arange(self.nF), self.f_dir]
        self.f_bnd = (nm == 0) | (nm == n)
        self.c_bnd = np.zeros(self.nC, bool)


_GRIDS: dict = {}


def fine_grid(n: int) -> FineGrid:
    if n not in _GRIDS:
        _GRIDS[n] = FineGrid(n)
    return _GRIDS[n]


# --------------------------------------------------------------------------
# Reference-cell shape functions at the 8 Gauss points
# --------------------------------------------------------------------------

_G = (0.5 - 0.5 / math.sqrt(3.0), 0.5 + 0.5 / math.sqrt(3.0))
QP = np.array([[_G[q & 1], _G[(q >> 1) & 1], _G[q >> 2]] for q in range(8)])   # x fastest

# Nedelec line table: (direction d, transverse corner bits in the order of the two other axes)
#   line -> (d, {axis: bit})
_LINES = [
    (1, {0: 0, 2: 0}), (1, {0: 1, 2: 0}), (0, {1: 0, 2: 0}), (0, {1: 1, 2: 0}),
    (1, {0: 0, 2: 1}), (1, {0: 1, 2: 1}), (0, {1: 0, 2: 1}), (0, {1: 1, 2: 1}),
    (2, {0: 0, 1: 0}), (2, {0: 1, 1: 0}), (2, {0: 0, 1: 1}), (2, {0: 1, 1: 1}),
]
_FACES = [(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)]   # face -> (normal axis, side)


def _w(bit, s):
    return s if bit else 1.0 - s


def _dw(bit):
    return 1.0 if bit else -1.0


def q1_ref(xi):
    """values[..., 8], reference gradients[..., 8, 3] at xi[..., 3]"""
    val = np.empty(xi.shape[:-1] + (8,)); grad = np.empty(xi.shape[:-1] + (8, 3))
    for v in range(8):
        b = (v & 1, (v >> 1) & 1, v >> 2)
        w = [_w(b[d], xi[..., d]) for d in range(3)]
        val[..., v] = w[0] * w[1] * w[2]
        grad[..., v, 0] = _dw(b[0]) * w[1] * w[2]
        grad[..., v, 1] = w[0] * _dw(b[1]) * w[2]
        grad[..., v, 2] = w[0] * w[1] * _dw(b[2])
    return val, grad


def ned_ref(xi):
    """reference values[..., 12, 3] and reference curls[..., 12, 3] (unit cube)"""
    val = np.zeros(xi.shape[:-1] + (12, 3)); curl = np.zeros(xi.shape[:-1] + (12, 3))
    for l, (d, tb) in enumerate(_LINES):
        (a0, b0), (a1, b1) = sorted(tb.items())
        f = _w(b0, xi[..., a0]) * _w(b1, xi[..., a1])
        gradf = np.zeros(xi.shape[:-1] + (3,))
        gradf[..., a0] = _dw(b0) * _w(b1, xi[..., a1])
        gradf[..., a1] = _w(b0, xi[..., a0]) * _dw(b1)
        val[..., l, d] = f
        e = np.zeros(3); e[d] = 1.0
        curl[..., l, :] = np.cross(gradf, e)
    return val, curl


def rt_ref(xi):
    """reference values[..., 6, 3], reference divergences[..., 6]"""
    val = np.zeros(xi.shape[:-1] + (6, 3)); div = np.zeros(xi.shape[:-1] + (6,))
    for f, (d, s) in enumerate(_FACES):
        val[..., f, d] = _w(s, xi[..., d])
        div[..., f] = _dw(s)
    return val, div


# Coarse standard shape functions at physical points (cube K = x0 + [0,H]^3)
def coarse_q1(x0, H, pts):
    v, g = q1_ref((pts - x0) / H)
    return v, g / H


def coarse_ned(x0, H, pts):
    v, c = ned_ref((pts - x0) / H)
    return v / H, c / H ** 2


def coarse_rt(x0, H, pts):
    v, d = rt_ref((pts - x0) / H)
    return v / H ** 2, d / H ** 3


# --------------------------------------------------------------------------
# Coefficients sampled at physical quadrature points
# --------------------------------------------------------------------------


def coefficient_fields(prob: Problem, x0, H, cell_gid: int = 0):
    """Returns (A[nC,8,3,3], Ainv[nC,8,3,3], B[nC,8], pts[nC,8,3]) for one coarse cell.

    Sine family: eqn_coeff_A.cc:162-242, eqn_coeff_B.cc:72-91.  Random field:
    BASELINE.md section 3 (piecewise constant per fine cell, keyed by the global
    fine-cell index cell_gid * n^3 + local index).
    """
    g = fine_grid(prob.n)
    h = H / prob.n
    pts = np.asarray(x0)[None, None, :] + h * (g.cell_ijk[:, None, :] + QP[None, :, :])
    R = rotation_matrix(prob.a_rotate)
    if prob.random_field_seed:
        gidx = np.uint64(cell_gid) * np.uint64(g.nC) + np.arange(g.nC, dtype=np.uint64)
        xi = random_field_normals(prob.random_field_seed, gidx)
        diag = np.exp(prob.random_field_sigma * xi[:, :3])[:, None, :].repeat(8, 1)
        B = np.exp(prob.random_field_sigma * xi[:, 3])[:, None].repeat(8, 1)
    else:
        diag = np.stack([prob.a_scale[d] * (1.0 - prob.a_alpha[d] * np.sin(2 * math.pi * prob.a_freq[d] * pts[..., d]))
                         for d in range(3)], -1)
        bexpr = Expr(prob.b_expr, {"frequency": prob.b_freq, "scale": prob.b_scale, "alpha": prob.b_alpha})
        B = bexpr(pts)
    A = np.einsum("ia,cqa,ja->cqij", R, diag, R)
    Ainv = np.einsum("ia,cqa,ja->cqij", R, 1.0 / diag, R)
    return A, Ainv, B, pts


# --------------------------------------------------------------------------
# Assembly of the (unconstrained) block matrices on the shared pattern
# --------------------------------------------------------------------------


def _scatter(rows_l2g, cols_l2g, local, nrow, ncol):
    nc, a, b = local.shape
    r = np.repeat(rows_l2g[:, :, None], b, 2).ravel()
    c = np.repeat(cols_l2g[:, None, :], a, 1).ravel()
    return sp.coo_matrix((local.ravel(), (r, c)), shape=(nrow, ncol)).tocsr()


@dataclass
class CellSystem:
    pairing: str
    n: int
    H: float
    x0: np.ndarray
    A00: sp.csr_matrix           # sigma-sigma (or the single SPD block for Q)
    K: sp.csr_matrix | None      # block(1,0); block(0,1) = -K^T
    A11: sp.csr_matrix | None
    global_rhs: np.ndarray       # on the u-type DoFs (all DoFs for Q)
    G0: np.ndarray               # [k, N0] essential boundary data (zero on interior)
    G1: np.ndarray | None        # [k, N1]
    F1: np.ndarray | None        # [k, N1] basis-specific volume rhs on u-type DoFs
    bnd0: np.ndarray
    bnd1: np.ndarray | None
    k0: int                      # number of sigma-type coarse functions
    k1: int


def assemble_cell(prob: Problem, corners: np.ndarray, cell_gid: int = 0) -> CellSystem:
    """corners[8,3] in deal.II vertex order; must be an axis-aligned cube."""
    corners = np.asarray(corners, float)
    x0 = corners[0]
    H = corners[7, 0] - x0[0]
    ref = x0[None, :] + H * np.array([[v & 1, (v >> 1) & 1, v >> 2] for v in range(8)], float)
    if not np.allclose(corners, ref, rtol=0, atol=1e-12 * max(1.0, abs(H))):
        raise ValueError("coarse cell is not an axis-aligned cube")
    n = prob.n
    g = fine_grid(n)
    h = H / n
    JxW = h ** 3 / 8.0
    A, Ainv, B, pts = coefficient_fields(prob, x0, H, cell_gid)
    q1v, q1g = q1_ref(QP); q1g = q1g / h                     # [8q,8], [8q,8,3]
    nedv, nedc = ned_ref(QP); nedv = nedv / h; nedc = nedc / h ** 2
    rtv, rtd = rt_ref(QP); rtv = rtv / h ** 2; rtd = rtd / h ** 3
    p = prob.pairing
    rhs_exprs = [Expr(s, prob.rhs_constants) for s in prob.rhs_expr.split(";")]

    def rhs_vec():
        if len(rhs_exprs) != 3:
            raise ValueError("vector right-hand side needs 3 ';'-separated components")
        return np.stack([e(pts) for e in rhs_exprs], -1)          # [nC,8,3]

    if p == "Q":
        # q_basis.cc:228-236
        loc = np.einsum("qia,cqab,qjb->cij", q1g, A, q1g) * JxW
        A00 = _scatter(g.cV, g.cV, loc, g.nV, g.nV)
        f = rhs_exprs[0](pts)
        lr = np.einsum("qi,cq->ci", q1v, f) * JxW
        grhs = np.bincount(g.cV.ravel(), lr.ravel(), g.nV)
        cv, _ = coarse_q1(x0, H, x0 + h * g.v_pos)
        G0 = np.where(g.v_bnd[None, :], cv.T, 0.0)
        return CellSystem(p, n, H, x0, A00, None, None, grhs, G0, None, None, g.v_bnd, None, 8, 0)

    if p == "Q_NED":
        # q_ned_basis.cc:483-505 : local = [Q1 8 | Ned 12]
        Binv = 1.0 / B
        l00 = np.einsum("qi,cq,qj->cij", q1v, Binv, q1v) * JxW
        l10 = np.einsum("qia,qja->ij", nedv, q1g)[None].repeat(g.nC, 0) * JxW      # (v_i, grad sigma_j)
        l11 = np.einsum("qia,cqab,qjb->cij", nedc, A, nedc) * JxW
        A00 = _scatter(g.cV, g.cV, l00, g.nV, g.nV)
        K = _scatter(g.cE, g.cV, l10, g.nE, g.nV)
        A11 = _scatter(g.cE, g.cE, l11, g.nE, g.nE)
        lr = np.einsum("qia,cqa->ci", nedv, rhs_vec()) * JxW
        grhs = np.bincount(g.cE.ravel(), lr.ravel(), g.nE)
        cv, cg = coarse_q1(x0, H, x0 + h * g.v_pos)                 # [nV,8], [nV,8,3]
        _, ceg = coarse_q1(x0, H, x0 + h * g.e_pos)                 # grad Q1 at edge midpoints
        cn, _ = coarse_ned(x0, H, x0 + h * g.e_pos)                 # [nE,12,3]
        ar = np.arange(g.nE)
        G0 = np.zeros((20, g.nV)); G1 = np.zeros((20, g.nE)); F1 = np.zeros((20, g.nE))
        G0[:8] = np.where(g.v_bnd[None, :], cv.T, 0.0)
        G1[:8] = np.where(g.e_bnd[None, :], h * ceg[ar, :, g.e_dir].T, 0.0)
        G1[8:] = np.where(g.e_bnd[None, :], h * cn[ar, :, g.e_dir].T, 0.0)
        _, cgq = coarse_q1(x0, H, pts)                              # [nC,8q,8m,3]
        lf = np.einsum("qia,cqma->cmi", nedv, cgq) * JxW            # (v_i, grad Q1_m)
        for m in range(8):
            F1[m] = np.bincount(g.cE.ravel(), lf[:, m, :].ravel(), g.nE)
        return CellSystem(p, n, H, x0, A00, K, A11, grhs, G0, G1, F1, g.v_bnd, g.e_bnd, 8, 12)

    if p == "NED_RT":
        # ned_rt_basis.cc:488-511 : local = [Ned 12 | RT 6]
        l00 = np.einsum("qia,cqab,qjb->cij", nedv, Ainv, nedv) * JxW
        l10 = np.einsum("qia,qja->ij", rtv, nedc)[None].repeat(g.nC, 0) * JxW      # (v_i, curl sigma_j)
        l11 = np.einsum("qi,cq,qj->cij", rtd, B, rtd) * JxW
        A00 = _scatter(g.cE, g.cE, l00, g.nE, g.nE)
        K = _scatter(g.cF, g.cE, l10, g.nF, g.nE)
        A11 = _scatter(g.cF, g.cF, l11, g.nF, g.nF)
        lr = np.einsum("qia,cqa->ci", rtv, rhs_vec()) * JxW
        grhs = np.bincount(g.cF.ravel(), lr.ravel(), g.nF)
        cn, _ = coarse_ned(x0, H, x0 + h * g.e_pos)
        cr, _ = coarse_rt(x0, H, x0 + h * g.f_pos)
        G0 = np.zeros((18, g.nE)); G1 = np.zeros((18, g.nF)); F1 = np.zeros((18, g.nF))
        G0[:12] = np.where(g.e_bnd[None, :], h * cn[np.arange(g.nE), :, g.e_dir].T, 0.0)
        G1[12:] = np.where(g.f_bnd[None, :], h * h * cr[np.arange(g.nF), :, g.f_dir].T, 0.0)
        _, ccq = coarse_ned(x0, H, pts)                             # curls [nC,8q,12m,3]
        lf = np.einsum("qia,cqma->cmi", rtv, ccq) * JxW
        for m in range(12):
            F1[m] = np.bincount(g.cF.ravel(), lf[:, m, :].ravel(), g.nF)
        return CellSystem(p, n, H, x0, A00, K, A11, grhs, G0, G1, F1, g.e_bnd, g.f_bnd, 12, 6)

    if p == "RT_DQ":
        # rt_dq_basis.cc:495-537 : local = [RT 6 | DG 1]
        l00 = np.einsum("qia,cqab,qjb->cij", rtv, Ainv, rtv) * JxW
        l10 = (rtd.sum(0) * JxW)[None, None, :].repeat(g.nC, 0)                    # (1, div psi_j)
        A00 = _scatter(g.cF, g.cF, l00, g.nF, g.nF)
        K = _scatter(g.cC, g.cF, l10, g.nC, g.nF)
        A11 = sp.csr_matrix((g.nC, g.nC))
        f = rhs_exprs[0](pts)
        grhs = f.sum(1) * JxW
        cr, _ = coarse_rt(x0, H, x0 + h * g.f_pos)
        G0 = np.where(g.f_bnd[None, :], h * h * cr[np.arange(g.nF), :, g.f_dir].T, 0.0)
        G1 = np.zeros((6, g.nC)); F1 = np.zeros((6, g.nC))
        for m in range(6):
            F1[m] = (1.0 if m & 1 else -1.0) * h ** 3 / H ** 3
        return CellSystem(p, n, H, x0, A00, K, A11, grhs, G0, G1, F1, g.f_bnd, g.c_bnd, 6, 0)

    raise ValueError(f"unknown pairing {p}")


# --------------------------------------------------------------------------
# Exact basis solve (== condense + SparseDirectUMFPACK + distribute) and Gram
# --------------------------------------------------------------------------


def solve_basis(cs: CellSystem):
    """Returns (X0[k,N0], X1[k,N1] or None): fine-scale basis functions."""
    i0 = np.flatnonzero(~cs.bnd0)
    k = cs.G0.shape[0]
    if cs.pairing == "Q":
        Aii = cs.A00[i0][:, i0].tocsc()
        rhs = -(cs.A00 @ cs.G0.T)[i0]
        X0 = cs.G0.copy()
        X0[:, i0] = spla.splu(Aii).solve(rhs).T
        return X0, None
    i1 = np.flatnonzero(~cs.bnd1)
    K = cs.K
    f0 = -(cs.A00 @ cs.G0.T) + K.T @ cs.G1.T                     # row 0: A00 s - K^T u = 0
    f1 = cs.F1.T - K @ cs.G0.T - cs.A11 @ cs.G1.T                # row 1: K s + A11 u = F1
    A00 = cs.A00[i0][:, i0]; K_ = K[i1][:, i0]; A11 = cs.A11[i1][:, i1]
    rhs = np.concatenate([f0[i0], f1[i1]])
    if cs.pairing == "RT_DQ":
        # constant-u null space (rt_dq_basis.cc:683-709): pin the first u DoF; sigma is unique
        keep = np.arange(1, len(i1))
        K_ = K_[keep]; A11 = A11[keep][:, keep]
        rhs = np.concatenate([f0[i0], f1[i1][keep]])
        i1 = i1[keep]
    S = sp.bmat([[A00, -K_.T], [K_, A11]], format="csc")
    sol = spla.splu(S).solve(rhs)
    X0 = cs.G0.copy(); X1 = cs.G1.copy()
    X0[:, i0] = sol[: len(i0)].T
    X1[:, i1] = sol[len(i0):].T
    return X0, X1


def element_matrix(cs: CellSystem, X0, X1):
    """Coarse element matrix M[k,k] (row-major, sigma-type first) and rhs r[k]."""
    if cs.pairing == "Q":
        M = X0 @ (cs.A00 @ X0.T)
        r = X0 @ cs.global_rhs
        return M, r
    if cs.pairing == "RT_DQ":
        S = X0                                                     # sigma parts of the 6 bases
        one = np.ones((1, cs.K.shape[0]))                          # u := 1 (set_u_to_std)
        M = np.zeros((7, 7)); r = np.zeros(7)
        M[:6, :6] = S @ (cs.A00 @ S.T)
        M[:6, 6:] = -(S @ (cs.K.T @ one.T))
        M[6:, :6] = one @ (cs.K @ S.T)
        M[6, 6] = (one @ (cs.A11 @ one.T))[0, 0]
        r[6] = one[0] @ cs.global_rhs
        return M, r
    k0, k1 = cs.k0, cs.k1
    S = X0[:k0]; U = X1[k0:]
    M = np.zeros((k0 + k1, k0 + k1)); r = np.zeros(k0 + k1)
    M[:k0, :k0] = S @ (cs.A00 @ S.T)
    M[:k0, k0:] = -(S @ (cs.K.T @ U.T))
    M[k0:, :k0] = U @ (cs.K @ S.T)
    M[k0:, k0:] = U @ (cs.A11 @ U.T)
    r[k0:] = U @ cs.global_rhs
    return M, r


def build_basis(prob: Problem, corners: np.ndarray, cell_gid: int = 0):
    cs = assemble_cell(prob, corners, cell_gid)
    X0, X1 = solve_basis(cs)
    M, r = element_matrix(cs, X0, X1)
    return M, r, X0, X1, cs


def k_of(pairing: str) -> int:
    return {"Q": 8, "Q_NED": 20, "NED_RT": 18, "RT_DQ": 7}[pairing]


# --------------------------------------------------------------------------
# Coarse mesh enumeration: p4est Morton (z-curve) order of the uniformly refined
# unit cube (ned_rt_global.cc:39-46, 61-63)
# --------------------------------------------------------------------------


def morton_cells(g_ref: int) -> np.ndarray:
    """corners[8^g, 8, 3] of the coarse cells in p4est z-order."""
    m = 1 << g_ref
    idx = np.arange(m ** 3, dtype=np.int64)
    ijk = np.zeros((m ** 3, 3), dtype=np.int64)
    for b in range(g_ref):
        for d in range(3):
            ijk[:, d] |= ((idx >> (3 * b + d)) & 1) << b
    H = 1.0 / m
    x0 = ijk * H
    off = np.array([[v & 1, (v >> 1) & 1, v >> 2] for v in range(8)], float) * H
    return x0[:, None, :] + off[None, :, :]


def partition(n_cells: int, rank: int, world: int):
    """Contiguous Morton chunk owned by rank (is_locally_owned equivalent)."""
    return (rank * n_cells) // world, ((rank + 1) * n_cells) // world


# --------------------------------------------------------------------------
# Nedelec L2 projection (pins the DoF convention against the reference's test)
# --------------------------------------------------------------------------


def project_constant_on_nedelec(n: int, vec) -> np.ndarray:
    """L2-project a constant vector field on FE_Nedelec(0) over [0,1]^3 with n^3 cells.

    Mirrors test/test_fe_projection_nedelec_mpi.cc:112-117 (MyVectorTools::project_on_fe_space).
    """
    g = fine_grid(n)
    h = 1.0 / n
    nedv, _ = ned_ref(QP); nedv = nedv / h
    JxW = h ** 3 / 8
    loc = np.einsum("qia,qja->ij", nedv, nedv)[None].repeat(g.nC, 0) * JxW
    Mm = _scatter(g.cE, g.cE, loc, g.nE, g.nE).tocsc()
    lr = np.einsum("qia,a->i", nedv, np.asarray(vec, float)) * JxW
    rhs = np.bincount(g.cE.ravel(), np.tile(lr, g.nC), g.nE)
    return spla.splu(Mm).solve(rhs)
