"""Integer-coded sampler handle over the C ABI (one model = one network x n_datasets x n_graphs chains).

This is the layer ``ModelORNOR`` (model.py) sits on after it has mapped symbols to integers; the
parity tests drive it directly with the same integer problem they hand to the oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import GbnError, KERNEL_AUTO, KERNEL_FAST, KERNEL_STRICT  # noqa: F401

DEFAULT_SPRIOR = (0.40, 0.40, 0.20, 0.30, 0.40, 0.30, 0.20, 0.40, 0.40)   # reference GraphORNOR.h:31-33


def device_count() -> int:
    return int(_capi.load().gbn_device_count())


class Sampler:
    """Chains of the OR-NOR network resident on one GPU.

    Parameters mirror gbn::ModelORNOR's ctor (reference libgbnet/include/ModelORNOR.h:26-29).
    ``evid_val`` may be 2-D ``[n_datasets, n_evid]`` to batch contrasts that share the network.
    """

    def __init__(self, src, trg, mor, evid_trg, evid_val, active_tf=(), sprior=DEFAULT_SPRIOR,
                 z_alpha=25., z_beta=25., z0_alpha=None, z0_beta=None, t_alpha=2., t_beta=2.,
                 n_graphs=3, noise_listen_children=True, comp_yprob=False, const_params=False,
                 t_focus=2., t_lmargin=2., t_hmargin=2., zn_focus=2., zn_lmargin=8., zn_hmargin=2.,
                 seed=0, device=0, kernel=KERNEL_AUTO, chain_offset=0, n_graphs_local=0):
        self._lib = _capi.load()
        self._h = None
        z0_alpha = z_alpha if z0_alpha is None else z0_alpha
        z0_beta = z_beta if z0_beta is None else z0_beta
        src = np.ascontiguousarray(src, np.uint32)
        trg = np.ascontiguousarray(trg, np.uint32)
        mor = np.ascontiguousarray(mor, np.int32)
        if not (src.shape == trg.shape == mor.shape and src.ndim == 1):
            raise ValueError("src, trg, mor must be 1-D arrays of equal length")
        evid_trg = np.ascontiguousarray(evid_trg, np.uint32)
        evid_val = np.ascontiguousarray(evid_val, np.int32)
        if evid_val.ndim == 1:
            evid_val = evid_val[None, :]
        if evid_val.shape[1] != evid_trg.size:
            raise ValueError("evid_val must have one column per evid_trg entry")
        act = np.ascontiguousarray(sorted(set(int(a) for a in active_tf)), np.uint32)
        net = _capi.Network(src.size, _capi.u32ptr(src), _capi.u32ptr(trg), _capi.i32ptr(mor),
                            act.size, _capi.u32ptr(act))
        ev = _capi.Evidence(evid_val.shape[0], evid_trg.size, _capi.u32ptr(evid_trg), _capi.i32ptr(evid_val))
        p = _capi.Params()
        self._lib.gbn_params_default(C.byref(p))
        sp = np.asarray(sprior, np.float64).ravel()
        if sp.size != 9:
            raise ValueError("sprior must have 9 entries")
        for i in range(9):
            p.sprior[i] = float(sp[i])
        p.z_alpha, p.z_beta, p.z0_alpha, p.z0_beta = float(z_alpha), float(z_beta), float(z0_alpha), float(z0_beta)
        p.t_alpha, p.t_beta = float(t_alpha), float(t_beta)
        p.n_graphs = int(n_graphs)
        p.noise_listen_children = int(bool(noise_listen_children))
        p.comp_yprob = int(bool(comp_yprob))
        p.const_params = int(bool(const_params))
        p.t_focus, p.t_lmargin, p.t_hmargin = float(t_focus), float(t_lmargin), float(t_hmargin)
        p.zn_focus, p.zn_lmargin, p.zn_hmargin = float(zn_focus), float(zn_lmargin), float(zn_hmargin)
        p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        p.device = int(device)
        p.kernel = int(kernel)
        p.chain_offset = int(chain_offset)
        p.n_graphs_local = int(n_graphs_local)
        h = C.c_void_p()
        _capi.check(self._lib.gbn_model_create(C.byref(net), C.byref(ev), C.byref(p), C.byref(h)))
        self._h = h
        d = _capi.Dims()
        _capi.check(self._lib.gbn_model_dims(self._h, C.byref(d)))
        self.dims = d
        self.n_evid = int(evid_trg.size)

    # ---------------------------------------------------------------- life cycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.gbn_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- dimensions
    def n_tf(self): return int(self.dims.n_tf)
    def n_targets(self): return int(self.dims.n_trg)
    def n_edges(self): return int(self.dims.n_edge)
    def n_nodes(self): return int(self.dims.n_nodes)
    def n_stats(self): return int(self.dims.n_stats)
    def n_chains(self): return int(self.dims.n_datasets * self.dims.n_graphs_local)
    def n_datasets(self): return int(self.dims.n_datasets)
    def kernel(self): return int(self._lib.gbn_model_kernel(self._h))
    def path(self):
        """0 strict kernel, 1 fast multi-kernel path (state in HBM), 2 fast persistent kernel (state in shared memory)"""
        return int(self._lib.gbn_model_path(self._h))
    def simulation(self): return bool(self.dims.simulation)

    def tf_ids(self):
        out = np.empty(self.n_tf(), np.uint32)
        _capi.check(self._lib.gbn_model_tf_ids(self._h, _capi.u32ptr(out)))
        return out

    def trg_ids(self):
        out = np.empty(self.n_targets(), np.uint32)
        _capi.check(self._lib.gbn_model_trg_ids(self._h, _capi.u32ptr(out)))
        return out

    # ---------------------------------------------------------------- sampling (reference ModelBase API)
    def sample(self, N=50000, dN=30):
        _capi.check(self._lib.gbn_model_sample(self._h, int(N), int(dN)))

    def sample_n(self, n):
        _capi.check(self._lib.gbn_model_sample_n(self._h, int(n)))

    def run_protocol(self, burn_its=10, max_its=100, N=20, dN=30):
        """The burn-in / convergence schedule of the reference's sample_model (gbnet/utils.py:93-135) as ONE
        library call: returns (calls made, last max R-hat)."""
        its, rhat = C.c_uint32(0), C.c_double(0.)
        _capi.check(self._lib.gbn_model_run_protocol(self._h, int(burn_its), int(max_its), int(N), int(dN),
                                                     C.byref(its), C.byref(rhat)))
        return int(its.value), float(rhat.value)

    def burn_stats(self):
        _capi.check(self._lib.gbn_model_burn_stats(self._h))

    def chain_length(self):
        v = C.c_double()
        _capi.check(self._lib.gbn_model_chain_length(self._h, C.byref(v)))
        return v.value

    def max_gelman_rubin(self):
        v = C.c_double()
        _capi.check(self._lib.gbn_model_max_gelman_rubin(self._h, C.byref(v)))
        return v.value

    def gelman_rubin(self, dataset=0):
        out = np.empty(self.n_nodes(), np.float64)
        _capi.check(self._lib.gbn_model_gelman_rubin(self._h, int(dataset), _capi.dptr(out)))
        return out

    def _posterior(self, fn, name, dataset):
        kind = name.encode()[:1] if len(name) == 1 else b"?"
        n = C.c_uint32()
        _capi.check(fn(self._h, int(dataset), kind, None, C.byref(n)))
        out = np.empty(n.value, np.float64)
        if n.value:
            _capi.check(fn(self._h, int(dataset), kind, _capi.dptr(out), C.byref(n)))
        return out

    def posterior_means(self, name, dataset=0):
        return self._posterior(self._lib.gbn_model_posterior_means, name, dataset)

    def posterior_sdevs(self, name, dataset=0):
        return self._posterior(self._lib.gbn_model_posterior_sdevs, name, dataset)

    def print_stats(self):
        _capi.check(self._lib.gbn_model_print_stats(self._h))

    def set_evidence(self, evid_trg, evid_val):
        evid_trg = np.ascontiguousarray(evid_trg, np.uint32)
        evid_val = np.ascontiguousarray(evid_val, np.int32)
        if evid_val.ndim == 1:
            evid_val = evid_val[None, :]
        ev = _capi.Evidence(evid_val.shape[0], evid_trg.size, _capi.u32ptr(evid_trg), _capi.i32ptr(evid_val))
        _capi.check(self._lib.gbn_model_set_evidence(self._h, C.byref(ev)))

    # ---------------------------------------------------------------- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(_capi.COMM_ID_BYTES)
        _capi.check(_capi.load().gbn_comm_unique_id(buf))
        return buf.raw

    def attach_comm(self, rank, world, uid: bytes):
        buf = C.create_string_buffer(uid, _capi.COMM_ID_BYTES)
        _capi.check(self._lib.gbn_model_attach_comm(self._h, int(rank), int(world), buf))

    # ---------------------------------------------------------------- parity hooks
    def get_state(self, chain):
        out = np.empty(self.n_nodes(), np.float64)
        _capi.check(self._lib.gbn_model_get_state(self._h, int(chain), _capi.dptr(out)))
        return out

    def set_state(self, chain, state):
        state = np.ascontiguousarray(state, np.float64)
        if state.size != self.n_nodes():
            raise ValueError("state has the wrong length")
        _capi.check(self._lib.gbn_model_set_state(self._h, int(chain), _capi.dptr(state)))

    def all_conditionals(self, chain):
        out = np.empty((self.n_nodes(), 3), np.float64)
        _capi.check(self._lib.gbn_model_eval_conditionals(self._h, int(chain), _capi.dptr(out)))
        return out

    def target_likelihoods(self, chain):
        out = np.empty(self.n_targets(), np.float64)
        _capi.check(self._lib.gbn_model_target_likelihoods(self._h, int(chain), _capi.dptr(out)))
        return out

    def joint_logprob(self, chain):
        v = C.c_double()
        _capi.check(self._lib.gbn_model_joint_logprob(self._h, int(chain), C.byref(v)))
        return v.value

    def inject(self, flat_u, beta_v, exp_v):
        """Draws for the next sweeps of EVERY local chain: flat_u[chain, sweep, nX+E],
        beta_v / exp_v [chain, sweep, nX]."""
        f = np.ascontiguousarray(flat_u, np.float64)
        b = np.ascontiguousarray(beta_v, np.float64)
        e = np.ascontiguousarray(exp_v, np.float64)
        C_, nX, E = self.n_chains(), self.n_tf(), self.n_edges()
        per = 2 * self.n_targets() if self.simulation() else nX + E
        if f.ndim != 3 or f.shape[0] != C_ or f.shape[2] != per:
            raise ValueError("flat_u must be [n_chains, n_sweeps, n_tf + n_edge] (simulation mode: 2 n_targets)")
        n = f.shape[1]
        if b.shape != (C_, n, nX) or e.shape != (C_, n, nX):
            raise ValueError("beta_v / exp_v must be [n_chains, n_sweeps, n_tf]")
        _capi.check(self._lib.gbn_model_inject_draws(self._h, n, _capi.dptr(f), _capi.dptr(b), _capi.dptr(e)))

    def clear_injection(self):
        _capi.check(self._lib.gbn_model_inject_draws(self._h, 0, None, None, None))

    def export_draws(self, chain, sweep0, n):
        nX, E = self.n_tf(), self.n_edges()
        f = np.empty((n, 2 * self.n_targets() if self.simulation() else nX + E), np.float64)
        b = np.empty((n, nX), np.float64)
        e = np.empty((n, nX), np.float64)
        _capi.check(self._lib.gbn_model_export_draws(self._h, int(chain), int(sweep0), int(n),
                                                     _capi.dptr(f), _capi.dptr(b), _capi.dptr(e)))
        return f, b, e

    def node_moments(self, chain):
        mean = np.empty(self.n_stats(), np.float64)
        var = np.empty(self.n_stats(), np.float64)
        _capi.check(self._lib.gbn_model_node_moments(self._h, int(chain), _capi.dptr(mean), _capi.dptr(var)))
        return mean, var

    # ---------------------------------------------------------------- checkpoint / signal
    def checkpoint(self) -> bytes:
        n = C.c_size_t()
        _capi.check(self._lib.gbn_model_checkpoint_size(self._h, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _capi.check(self._lib.gbn_model_checkpoint_save(self._h, buf, n.value))
        return buf.raw

    def restore(self, blob: bytes):
        buf = C.create_string_buffer(blob, len(blob))
        _capi.check(self._lib.gbn_model_checkpoint_load(self._h, buf, len(blob)))

    @staticmethod
    def set_signal(signum: int):
        """gbn::set_signal (ModelBase.cpp:211-214): a non-zero value stops sample() after the
        current batch; sample() clears it (ModelBase.cpp:126)."""
        _capi.load().gbn_set_signal(int(signum))

    # ---------------------------------------------------------------- measurement
    def last_device_ms(self):
        v = C.c_double()
        _capi.check(self._lib.gbn_model_last_device_ms(self._h, C.byref(v)))
        return v.value

    def set_phase_timing(self, on=True):
        _capi.check(self._lib.gbn_model_set_phase_timing(self._h, int(bool(on))))

    def set_stream_groups(self, groups=0):
        _capi.check(self._lib.gbn_model_set_stream_groups(self._h, int(groups)))

    def phase_ms(self):
        """(ms in the X, T, S kernels, sweeps covered) since set_phase_timing(True)"""
        ms = (C.c_double * 3)()
        n = C.c_uint64()
        _capi.check(self._lib.gbn_model_phase_ms(self._h, ms, C.byref(n)))
        return [float(v) for v in ms], int(n.value)

    def launch_count(self):
        v = C.c_uint64()
        _capi.check(self._lib.gbn_model_launch_count(self._h, C.byref(v)))
        return int(v.value)

    def sweep_count(self):
        v = C.c_uint64()
        _capi.check(self._lib.gbn_model_sweep_count(self._h, C.byref(v)))
        return int(v.value)

    def timer_start(self):
        _capi.check(self._lib.gbn_model_timer_start(self._h))

    def timer_stop(self):
        v = C.c_double()
        _capi.check(self._lib.gbn_model_timer_stop(self._h, C.byref(v)))
        return v.value

    def debug_counters(self, enable=True):
        out = (C.c_uint64 * 16)()
        _capi.check(self._lib.gbn_model_debug_counters(self._h, int(bool(enable)), out))
        return [int(v) for v in out]

    def synchronize(self):
        _capi.check(self._lib.gbn_model_synchronize(self._h))


EVID_ABSENT = 127


def fisher_counts(src, trg, mor, evid_trg, evid_val, device=0):
    """Per contrast and source TF the six edge counts behind the reference's get_tests (gbnet/utils.py:10-35),
    reduced on the device for all contrasts at once.  ``evid_val``: [n_datasets, n_evid] values -1 / 0 / +1 or
    EVID_ABSENT.  Returns (sorted unique source ids, counts[n_datasets, n_tf, 6])."""
    lib = _capi.load()
    src = np.ascontiguousarray(src, np.uint32)
    trg = np.ascontiguousarray(trg, np.uint32)
    mor = np.ascontiguousarray(mor, np.int32)
    evid_trg = np.ascontiguousarray(evid_trg, np.uint32)
    evid_val = np.ascontiguousarray(evid_val, np.int8)
    if evid_val.ndim == 1:
        evid_val = evid_val[None, :]
    if evid_val.shape[1] != evid_trg.size:
        raise ValueError("evid_val must have one column per evid_trg entry")
    net = _capi.Network(src.size, _capi.u32ptr(src), _capi.u32ptr(trg), _capi.i32ptr(mor), 0, None)
    n_tf = C.c_uint32(0)

    def call(ids, counts):
        st = lib.gbn_fisher_counts(int(device), C.byref(net), evid_val.shape[0], evid_trg.size, _capi.u32ptr(evid_trg),
                                   evid_val.ctypes.data_as(C.POINTER(C.c_int8)), C.byref(n_tf), ids, counts)
        if st != 0:
            raise GbnError(st, lib.gbn_fisher_last_error().decode())

    call(None, None)
    ids = np.empty(n_tf.value, np.uint32)
    counts = np.zeros((evid_val.shape[0], n_tf.value, 6), np.uint32)
    call(_capi.u32ptr(ids), _capi.u32ptr(counts))
    return ids, counts


def philox_kat(ctr, key, device=0):
    lib = _capi.load()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    _capi.check(lib.gbn_philox_kat(int(device), c, k, o))
    return [int(v) for v in o]
