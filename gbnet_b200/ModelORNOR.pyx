# distutils: language = c++
# cython: language_level=3
"""Cython binding of the C ABI: the reference's gbnet/ModelORNOR.pyx re-pointed at libgbnet_b200.so.

Same class name (``PyModelORNOR``), constructor keywords, methods and return shapes as the reference
(gbnet/ModelORNOR.pyx:10-130).  The pandas marshaling is the reference's (gbnet_b200.model.marshal_frames
restates pyx:18-66); what changes is the last line of ``__cinit__``: instead of ``new ModelORNOR(...)``
on string containers, integer arrays cross ``gbn_model_create``.  ``sample`` releases the GIL.
"""
import numpy as np

from gbnet_b200.model import marshal_frames, check_sprior

from libc.stdint cimport uint32_t, int32_t
from libc.string cimport memset


cdef class PyModelORNOR:
    cdef gbn_model *h
    cdef object _tf_symbols, _trg_symbols, _edges, _const_params, _simulation

    def __cinit__(self, ents, rels, evid=None, set active_tf_set_py=set(),
                  sprior=None, z_alpha=25., z_beta=25., z0_alpha=None, z0_beta=None, t_alpha=2., t_beta=2.,
                  unsigned int n_graphs=3, bint noise_listen_children=True, bint comp_yprob=False,
                  bint const_params=False, unsigned int x_mode=1,
                  double t_focus=2., double t_lmargin=2., double t_hmargin=2., double zn_focus=2.,
                  double zn_lmargin=8., double zn_hmargin=2., seed=0, int device=0, int kernel=0):
        self.h = NULL
        z0_alpha = z_alpha if z0_alpha is None else z0_alpha
        z0_beta = z_beta if z0_beta is None else z0_beta
        m = marshal_frames(ents, rels, evid, active_tf_set_py)
        sprior_py = check_sprior(sprior)
        self._tf_symbols, self._edges, self._const_params = m["tf_symbols"], m["edges"], bool(const_params)
        self._trg_symbols, self._simulation = m["trg_symbols"], False

        cdef uint32_t[::1] src = np.ascontiguousarray(m["src"], np.uint32)
        cdef uint32_t[::1] trg = np.ascontiguousarray(m["trg"], np.uint32)
        cdef int32_t[::1] mor = np.ascontiguousarray(m["mor"], np.int32)
        cdef uint32_t[::1] et = np.ascontiguousarray(m["evid_trg"], np.uint32)
        cdef int32_t[::1] ev = np.ascontiguousarray(m["evid_val"], np.int32)
        cdef uint32_t[::1] act = np.ascontiguousarray(m["active"], np.uint32)
        cdef gbn_network net
        cdef gbn_evidence evd
        cdef gbn_params p
        memset(&net, 0, sizeof(net))
        memset(&evd, 0, sizeof(evd))
        gbn_params_default(&p)
        net.n_edge = <uint32_t> src.shape[0]
        if src.shape[0] > 0:
            net.src = &src[0]; net.trg = &trg[0]; net.mor = &mor[0]
        net.n_active = <uint32_t> act.shape[0]
        if act.shape[0] > 0:
            net.active_tf = &act[0]
        evd.n_datasets = 1
        evd.n_evid = <uint32_t> et.shape[0]
        if et.shape[0] > 0:
            evd.evid_trg = &et[0]; evd.evid_val = &ev[0]
        cdef int i
        for i in range(9):
            p.sprior[i] = sprior_py[i]
        p.z_alpha = z_alpha; p.z_beta = z_beta; p.z0_alpha = z0_alpha; p.z0_beta = z0_beta
        p.t_alpha = t_alpha; p.t_beta = t_beta
        p.n_graphs = n_graphs
        p.noise_listen_children = noise_listen_children; p.comp_yprob = comp_yprob; p.const_params = const_params
        p.t_focus = t_focus; p.t_lmargin = t_lmargin; p.t_hmargin = t_hmargin
        p.zn_focus = zn_focus; p.zn_lmargin = zn_lmargin; p.zn_hmargin = zn_hmargin
        p.seed = <unsigned long long> (int(seed) & 0xFFFFFFFFFFFFFFFF)
        p.device = device; p.kernel = kernel
        cdef int rc = gbn_model_create(&net, &evd, &p, &self.h)
        if rc != 0:
            msg = gbn_last_error().decode()
            # the reference's ctor is `except +`: out_of_range -> IndexError, invalid_argument -> ValueError
            raise {-1: ValueError, -3: NotImplementedError}.get(rc, RuntimeError)(msg)
        cdef gbn_dims dims
        self._check(gbn_model_dims(self.h, &dims))
        self._simulation = dims.simulation != 0

    cdef _check(self, int rc):
        if rc != 0:
            raise RuntimeError(gbn_last_error().decode())

    def _node_ids(self):
        if self._simulation:          # evid=None: H and Y per target, then T (GraphORNOR.cpp:89-106)
            ids = [[k, t] for t in self._trg_symbols for k in ("H", "Y")]
            if not self._const_params:
                ids += [["T", s] for s in self._tf_symbols]
            return ids
        ids = [["X", s] for s in self._tf_symbols]
        if not self._const_params:
            ids += [["T", s] for s in self._tf_symbols]
        ids += [["S", f"{s}-->{t}"] for s, t in self._edges]
        return ids

    def get_gelman_rubin(self):
        cdef gbn_dims d
        self._check(gbn_model_dims(self.h, &d))
        cdef double[::1] out = np.empty(d.n_nodes, np.float64)
        self._check(gbn_model_gelman_rubin(self.h, 0, &out[0]))
        return [f"{name}_{uid}".split('_')[:2] + [float(v)] for (name, uid), v in zip(self._node_ids(), out)]

    def get_max_gelman_rubin(self):
        cdef double v = 0
        self._check(gbn_model_max_gelman_rubin(self.h, &v))
        return v

    def sample(self, unsigned int N=50000, unsigned int deltaN=30):
        cdef int rc
        with nogil:                      # the reference holds the GIL for the whole call (pyx:92-94)
            rc = gbn_model_sample(self.h, N, deltaN)
        self._check(rc)
        return

    def sample_n(self, unsigned int n):
        cdef int rc
        with nogil:
            rc = gbn_model_sample_n(self.h, n)
        self._check(rc)

    def run_protocol(self, unsigned int burn_its=10, unsigned int max_its=100, unsigned int N=20, unsigned int dN=30):
        """The schedule of gbnet/utils.py:93-135 (sample_model) as one library call: (calls made, last max R-hat)."""
        cdef int rc
        cdef uint32_t its = 0
        cdef double rhat = 0
        with nogil:
            rc = gbn_model_run_protocol(self.h, burn_its, max_its, N, dN, &its, &rhat)
        self._check(rc)
        return int(its), float(rhat)

    def path(self):
        return gbn_model_path(self.h)

    def print_stats(self):
        self._check(gbn_model_print_stats(self.h))

    def burn_stats(self):
        self._check(gbn_model_burn_stats(self.h))

    cdef _posterior(self, var_name, bint want_sd):
        per = {"X": 2, "T": 1, "S": 3, "H": 3, "Y": 3}.get(var_name)
        if per is None:
            return []
        cdef uint32_t n = 0
        cdef char kind = ord(var_name)
        if want_sd:
            self._check(gbn_model_posterior_sdevs(self.h, 0, kind, NULL, &n))
        else:
            self._check(gbn_model_posterior_means(self.h, 0, kind, NULL, &n))
        if n == 0:
            return []
        cdef double[::1] out = np.empty(n, np.float64)
        if want_sd:
            self._check(gbn_model_posterior_sdevs(self.h, 0, kind, &out[0], &n))
        else:
            self._check(gbn_model_posterior_means(self.h, 0, kind, &out[0], &n))
        ids = [i for i in self._node_ids() if i[0] == var_name]
        res = []
        for k, (name, uid) in enumerate(ids):
            for j in range(per):
                res.append(f"{name}_{uid}_{j}".split('_') + [float(out[k * per + j])])
        return res

    def get_posterior_means(self, var_name):
        return self._posterior(var_name, False)

    def get_posterior_sdevs(self, var_name):
        return self._posterior(var_name, True)

    def kernel(self):
        return gbn_model_kernel(self.h)

    def __dealloc__(self):
        if self.h is not NULL:
            gbn_model_destroy(self.h)
            self.h = NULL
