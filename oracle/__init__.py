"""Parity oracles for the OR-NOR Gibbs hot path -- TEST INFRASTRUCTURE ONLY.

Two CPU implementations behind one duck-typed Python interface:

* ``PortModel``  -- the plain-C restatement in ``ornor_oracle.c`` (always available; built by
  ``make -C oracle port``).
* ``RefModel``   -- the reference's own C++ (``/root/reference/libgbnet/src/*.cpp``) compiled
  UNMODIFIED against ``gsl_shim/`` into ``oracle/_ref/libgbnet_ref.so`` (``make -C oracle ref``;
  only buildable where ``/root/reference`` exists, but the built ``.so`` travels to the GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
package.  The product (``gbnet_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(_HERE, "libornor_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libgbnet_ref.so")

DEFAULT_SPRIOR = (0.40, 0.40, 0.20, 0.30, 0.40, 0.30, 0.20, 0.40, 0.40)  # GraphORNOR.h:31-33


def build(reference: str = "/root/reference") -> None:
    """Compile the C port and, when the reference sources are present, oracle/_ref."""
    subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL)
    if os.path.isdir(os.path.join(reference, "libgbnet", "src")):
        subprocess.run(["make", "-C", _HERE, "ref", f"REFERENCE={reference}"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


@dataclass
class Problem:
    """Integer-coded model input: what ModelORNOR's ctor receives after string marshaling
    (gbnet/ModelORNOR.pyx:56-83)."""
    src: np.ndarray                  # [E] TF uid per edge, edge input order
    trg: np.ndarray                  # [E] target uid per edge
    mor: np.ndarray                  # [E] -1/0/+1
    evid_trg: np.ndarray             # [n_evid]
    evid_val: np.ndarray             # [n_evid] -1/0/+1
    active_tf: np.ndarray = field(default_factory=lambda: np.zeros(0, np.uint32))
    sprior: tuple = DEFAULT_SPRIOR
    z_alpha: float = 25.0
    z_beta: float = 25.0
    z0_alpha: float = 25.0
    z0_beta: float = 25.0
    t_alpha: float = 2.0
    t_beta: float = 2.0
    n_graphs: int = 3
    noise_listen_children: bool = True
    comp_yprob: bool = False
    const_params: bool = False
    t_focus: float = 2.0
    t_lmargin: float = 2.0
    t_hmargin: float = 2.0
    zn_focus: float = 2.0
    zn_lmargin: float = 8.0
    zn_hmargin: float = 2.0
    seed: int = 12345

    def arrays(self):
        u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        return (u32(self.src), u32(self.trg), i32(self.mor), u32(self.evid_trg), i32(self.evid_val),
                u32(self.active_tf))


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class _Base:
    """Methods shared by both oracles; subclasses bind self._f(name) to the right symbol."""

    def n_chains(self): return int(self._f("n_chains")(self._h))
    def n_nodes(self): return int(self._f("n_nodes")(self._h))
    def n_targets(self): return int(self._f("n_targets")(self._h))

    def get_state(self, chain):
        out = np.empty(self.n_nodes(), np.float64)
        self._f("get_state")(self._h, chain, _ptr(out, C.c_double))
        return out

    def set_state(self, chain, state):
        state = np.ascontiguousarray(state, np.float64)
        assert state.size == self.n_nodes()
        self._f("set_state")(self._h, chain, _ptr(state, C.c_double))

    def blanket_loglik(self, chain, node, value):
        return float(self._f("blanket_loglik")(self._h, chain, node, float(value)))

    def all_conditionals(self, chain):
        out = np.empty((self.n_nodes(), 3), np.float64)
        self._f("all_conditionals")(self._h, chain, _ptr(out, C.c_double))
        return out

    def target_likelihoods(self, chain):
        out = np.empty(self.n_targets(), np.float64)
        self._f("target_likelihoods")(self._h, chain, _ptr(out, C.c_double))
        return out

    def joint_logprob(self, chain):
        return float(self._f("joint_logprob")(self._h, chain))

    def inject(self, chain, flat_u, beta_v, exp_v):
        f = np.ascontiguousarray(flat_u, np.float64).ravel()
        b = np.ascontiguousarray(beta_v, np.float64).ravel()
        e = np.ascontiguousarray(exp_v, np.float64).ravel()
        self._f("inject")(self._h, chain, _ptr(f, C.c_double), f.size, _ptr(b, C.c_double), b.size,
                          _ptr(e, C.c_double), e.size)

    def clear_injection(self, chain): self._f("clear_injection")(self._h, chain)
    def sample_n(self, n): self._f("sample_n")(self._h, n)
    def sample(self, N=50000, dN=30): self._f("sample")(self._h, N, dN)
    def burn_stats(self): self._f("burn_stats")(self._h)
    def time_sample_n(self, n): return float(self._f("time_sample_n")(self._h, n))
    def chain_length(self): return float(self._f("chain_length")(self._h))
    def node_mean(self, chain, node, stat): return float(self._f("node_mean")(self._h, chain, node, stat))
    def node_variance(self, chain, node, stat): return float(self._f("node_variance")(self._h, chain, node, stat))
    def max_gelman_rubin(self): return float(self._f("max_gelman_rubin")(self._h))

    def gelman_rubin(self):
        out = np.empty(self.n_nodes(), np.float64)
        self._f("gelman_rubin")(self._h, _ptr(out, C.c_double))
        return out

    def posterior_means(self, name):
        n = self._f("posterior_means")(self._h, name.encode(), None)
        out = np.empty(n, np.float64)
        self._f("posterior_means")(self._h, name.encode(), _ptr(out, C.c_double))
        return out

    def posterior_sdevs(self, name):
        n = self._f("posterior_sdevs")(self._h, name.encode(), None)
        out = np.empty(n, np.float64)
        self._f("posterior_sdevs")(self._h, name.encode(), _ptr(out, C.c_double))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._f("model_destroy")(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _decl(lib, name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


def _declare_common(lib, pfx):
    P = C.POINTER
    vp, u, d, sz = C.c_void_p, C.c_uint, C.c_double, C.c_size_t
    _decl(lib, pfx + "model_destroy", None, [vp])
    for n in ("n_chains", "n_nodes", "n_targets"):
        _decl(lib, pfx + n, u, [vp])
    _decl(lib, pfx + "get_state", None, [vp, u, P(d)])
    _decl(lib, pfx + "set_state", None, [vp, u, P(d)])
    _decl(lib, pfx + "blanket_loglik", d, [vp, u, u, d])
    _decl(lib, pfx + "all_conditionals", None, [vp, u, P(d)])
    _decl(lib, pfx + "target_likelihoods", None, [vp, u, P(d)])
    _decl(lib, pfx + "joint_logprob", d, [vp, u])
    _decl(lib, pfx + "inject", None, [vp, u, P(d), sz, P(d), sz, P(d), sz])
    _decl(lib, pfx + "clear_injection", None, [vp, u])
    _decl(lib, pfx + "sample_n", None, [vp, u])
    _decl(lib, pfx + "sample", None, [vp, u, u])
    _decl(lib, pfx + "burn_stats", None, [vp])
    _decl(lib, pfx + "time_sample_n", d, [vp, u])
    _decl(lib, pfx + "chain_length", d, [vp])
    _decl(lib, pfx + "node_mean", d, [vp, u, u, u])
    _decl(lib, pfx + "node_variance", d, [vp, u, u, u])
    _decl(lib, pfx + "max_gelman_rubin", d, [vp])
    _decl(lib, pfx + "gelman_rubin", None, [vp, P(d)])
    _decl(lib, pfx + "posterior_means", u, [vp, C.c_char_p, P(d)])
    _decl(lib, pfx + "posterior_sdevs", u, [vp, C.c_char_p, P(d)])
    _decl(lib, pfx + "omp_max_threads", C.c_int, [])
    _decl(lib, pfx + "omp_set_threads", None, [C.c_int])
    _decl(lib, pfx + "set_signal", None, [C.c_int])
    _decl(lib, pfx + "last_error", C.c_char_p, [])


class _OrnParams(C.Structure):
    _fields_ = [("sprior", C.c_double * 9),
                ("z_alpha", C.c_double), ("z_beta", C.c_double),
                ("z0_alpha", C.c_double), ("z0_beta", C.c_double),
                ("t_alpha", C.c_double), ("t_beta", C.c_double),
                ("n_graphs", C.c_uint),
                ("noise_listen_children", C.c_int), ("comp_yprob", C.c_int), ("const_params", C.c_int),
                ("t_focus", C.c_double), ("t_lmargin", C.c_double), ("t_hmargin", C.c_double),
                ("zn_focus", C.c_double), ("zn_lmargin", C.c_double), ("zn_hmargin", C.c_double),
                ("seed", C.c_uint64)]


_port_lib = None
_ref_lib = None


def port_lib():
    global _port_lib
    if _port_lib is None:
        if not os.path.exists(PORT_LIB):
            build()
        lib = C.CDLL(PORT_LIB)
        _declare_common(lib, "orn_")
        P = C.POINTER
        _decl(lib, "orn_model_create", C.c_void_p,
              [C.c_uint, P(C.c_uint), P(C.c_uint), P(C.c_int), C.c_uint, P(C.c_uint), P(C.c_int),
               C.c_uint, P(C.c_uint), P(_OrnParams)])
        _decl(lib, "orn_n_tf", C.c_uint, [C.c_void_p])
        _decl(lib, "orn_n_edges", C.c_uint, [C.c_void_p])
        _decl(lib, "orn_edge_slots", None, [C.c_void_p, P(C.c_uint)])
        _decl(lib, "orn_t_prior", None, [C.c_void_p, C.c_uint, P(C.c_double), P(C.c_double)])
        _decl(lib, "orn_export_draws", None,
              [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, P(C.c_double), P(C.c_double), P(C.c_double)])
        _decl(lib, "orn_philox4x32_10", None, [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)])
        _decl(lib, "orn_target_kat", None,
              [C.c_uint, P(C.c_double), P(C.c_uint), P(C.c_uint), C.c_double, P(C.c_double),
               P(C.c_double), P(C.c_double)])
        _decl(lib, "orn_beta_pdf", C.c_double, [C.c_double] * 3)
        _decl(lib, "orn_draw_beta", C.c_double, [C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_double])
        _decl(lib, "orn_draw_exp", C.c_double, [C.c_uint64, C.c_uint, C.c_uint, C.c_uint])
        _decl(lib, "orn_draw_u", C.c_double, [C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.c_uint])
        _port_lib = lib
    return _port_lib


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        if not os.path.exists(REF_LIB):
            raise FileNotFoundError(f"{REF_LIB} not built (needs /root/reference; run make -C oracle ref)")
        lib = C.CDLL(REF_LIB)
        _declare_common(lib, "ref_")
        P = C.POINTER
        _decl(lib, "ref_model_create", C.c_void_p,
              [C.c_uint, P(C.c_uint), P(C.c_uint), P(C.c_int), C.c_uint, P(C.c_uint), P(C.c_int),
               C.c_uint, P(C.c_uint), P(C.c_double)] + [C.c_double] * 6 +
              [C.c_uint, C.c_int, C.c_int, C.c_int] + [C.c_double] * 6)
        _decl(lib, "ref_node_info", C.c_uint, [C.c_void_p, C.c_uint, C.c_char_p, C.c_char_p, C.c_uint])
        _decl(lib, "ref_seed", None, [C.c_void_p, C.c_uint, C.c_ulong])
        _decl(lib, "ref_t_prior", None, [C.c_void_p, C.c_uint, P(C.c_double), P(C.c_double)])
        _ref_lib = lib
    return _ref_lib


class PortModel(_Base):
    """The plain-C restatement (ornor_oracle.c)."""

    def __init__(self, pb: Problem):
        self._lib = port_lib()
        src, trg, mor, et, ev, act = pb.arrays()
        prm = _OrnParams()
        for i, v in enumerate(pb.sprior):
            prm.sprior[i] = float(v)
        for name in ("z_alpha", "z_beta", "z0_alpha", "z0_beta", "t_alpha", "t_beta", "t_focus",
                     "t_lmargin", "t_hmargin", "zn_focus", "zn_lmargin", "zn_hmargin"):
            setattr(prm, name, float(getattr(pb, name)))
        prm.n_graphs = pb.n_graphs
        prm.noise_listen_children = int(pb.noise_listen_children)
        prm.comp_yprob = int(pb.comp_yprob)
        prm.const_params = int(pb.const_params)
        prm.seed = pb.seed
        self._keep = (src, trg, mor, et, ev, act)
        self._h = self._lib.orn_model_create(
            src.size, _ptr(src, C.c_uint), _ptr(trg, C.c_uint), _ptr(mor, C.c_int),
            et.size, _ptr(et, C.c_uint), _ptr(ev, C.c_int), act.size, _ptr(act, C.c_uint), C.byref(prm))
        if not self._h:
            raise ValueError(self._lib.orn_last_error().decode())
        self.seed = pb.seed

    def _f(self, name):
        return getattr(self._lib, "orn_" + name)

    def n_tf(self): return int(self._lib.orn_n_tf(self._h))
    def n_edges(self): return int(self._lib.orn_n_edges(self._h))

    def edge_slots(self):
        out = np.empty(self.n_edges(), np.uint32)
        self._lib.orn_edge_slots(self._h, _ptr(out, C.c_uint))
        return out

    def t_prior(self, tf):
        a, b = C.c_double(), C.c_double()
        self._lib.orn_t_prior(self._h, tf, C.byref(a), C.byref(b))
        return a.value, b.value

    def export_draws(self, chain, sweep0, n):
        nX, E = self.n_tf(), self.n_edges()
        flat = np.empty((n, nX + E), np.float64)
        beta = np.empty((n, nX), np.float64)
        expo = np.empty((n, nX), np.float64)
        self._lib.orn_export_draws(self._h, chain, sweep0, n, _ptr(flat, C.c_double),
                                   _ptr(beta, C.c_double), _ptr(expo, C.c_double))
        return flat, beta, expo


class RefModel(_Base):
    """The reference's own C++, compiled unmodified (oracle/_ref)."""

    def __init__(self, pb: Problem):
        self._lib = ref_lib()
        src, trg, mor, et, ev, act = pb.arrays()
        sp = np.asarray(pb.sprior, np.float64)
        self._h = self._lib.ref_model_create(
            src.size, _ptr(src, C.c_uint), _ptr(trg, C.c_uint), _ptr(mor, C.c_int),
            et.size, _ptr(et, C.c_uint), _ptr(ev, C.c_int), act.size, _ptr(act, C.c_uint),
            _ptr(sp, C.c_double),
            pb.z_alpha, pb.z_beta, pb.z0_alpha, pb.z0_beta, pb.t_alpha, pb.t_beta,
            pb.n_graphs, int(pb.noise_listen_children), int(pb.comp_yprob), int(pb.const_params),
            pb.t_focus, pb.t_lmargin, pb.t_hmargin, pb.zn_focus, pb.zn_lmargin, pb.zn_hmargin)
        if not self._h:
            raise ValueError(self._lib.ref_last_error().decode())

    def _f(self, name):
        return getattr(self._lib, "ref_" + name)

    def node_info(self, node):
        kind = C.create_string_buffer(2)
        ident = C.create_string_buffer(128)
        n_stats = self._lib.ref_node_info(self._h, node, kind, ident, 128)
        return kind.value.decode()[:1], ident.value.decode(), int(n_stats)

    def seed_chain(self, chain, seed): self._lib.ref_seed(self._h, chain, seed)

    def t_prior(self, node):
        a, b = C.c_double(), C.c_double()
        self._lib.ref_t_prior(self._h, node, C.byref(a), C.byref(b))
        return a.value, b.value


# ---------------------------------------------------------------- small helpers for KATs
def philox4x32_10(ctr, key):
    lib = port_lib()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.orn_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def target_kat(parents, z, yprob=(0.05, 0.90, 0.05)):
    """parents: list of (t, x, s).  Returns (P_model[3], L[3])."""
    lib = port_lib()
    n = len(parents)
    t = (C.c_double * max(n, 1))(*[p[0] for p in parents])
    x = (C.c_uint * max(n, 1))(*[p[1] for p in parents])
    s = (C.c_uint * max(n, 1))(*[p[2] for p in parents])
    yp = (C.c_double * 3)(*yprob)
    p = (C.c_double * 3)()
    l = (C.c_double * 3)()
    lib.orn_target_kat(n, t, x, s, z, yp, p, l)
    return list(p), list(l)
