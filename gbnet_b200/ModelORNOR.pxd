# C ABI of libgbnet_b200.so as seen by Cython: replaces the `cdef extern ... cppclass ModelORNOR`
# block of the reference (gbnet/ModelORNOR.pxd:24-39).
from libc.stdint cimport uint32_t, int32_t, uint64_t

cdef extern from "gbnet_b200.h":
    ctypedef struct gbn_model:
        pass
    ctypedef struct gbn_network:
        uint32_t n_edge
        const uint32_t *src
        const uint32_t *trg
        const int32_t *mor
        uint32_t n_active
        const uint32_t *active_tf
    ctypedef struct gbn_evidence:
        uint32_t n_datasets
        uint32_t n_evid
        const uint32_t *evid_trg
        const int32_t *evid_val
    ctypedef struct gbn_params:
        double sprior[9]
        double z_alpha
        double z_beta
        double z0_alpha
        double z0_beta
        double t_alpha
        double t_beta
        uint32_t n_graphs
        int32_t noise_listen_children
        int32_t comp_yprob
        int32_t const_params
        double t_focus
        double t_lmargin
        double t_hmargin
        double zn_focus
        double zn_lmargin
        double zn_hmargin
        uint64_t seed
        int32_t device
        int32_t kernel
        uint32_t chain_offset
        uint32_t n_graphs_local
    ctypedef struct gbn_dims:
        uint32_t n_tf
        uint32_t n_trg
        uint32_t n_edge
        uint32_t n_nodes
        uint32_t n_datasets
        uint32_t n_graphs
        uint32_t n_graphs_local
        uint32_t chain_offset
        uint32_t n_stats
        uint32_t simulation
    void gbn_params_default(gbn_params *p)
    const char *gbn_last_error()
    int gbn_model_create(const gbn_network *net, const gbn_evidence *ev, const gbn_params *p, gbn_model **out)
    void gbn_model_destroy(gbn_model *m)
    int gbn_model_dims(const gbn_model *m, gbn_dims *out)
    int gbn_model_kernel(const gbn_model *m)
    int gbn_model_sample(gbn_model *m, uint32_t N, uint32_t dN) nogil
    int gbn_model_sample_n(gbn_model *m, uint32_t n) nogil
    int gbn_model_run_protocol(gbn_model *m, uint32_t burn_its, uint32_t max_its, uint32_t N, uint32_t dN,
                               uint32_t *its_out, double *rhat_out) nogil
    int gbn_model_path(const gbn_model *m)
    int gbn_model_burn_stats(gbn_model *m)
    int gbn_model_max_gelman_rubin(gbn_model *m, double *out)
    int gbn_model_gelman_rubin(gbn_model *m, uint32_t dataset, double *out)
    int gbn_model_posterior_means(gbn_model *m, uint32_t dataset, char kind, double *out, uint32_t *n)
    int gbn_model_posterior_sdevs(gbn_model *m, uint32_t dataset, char kind, double *out, uint32_t *n)
    int gbn_model_print_stats(gbn_model *m)
    void gbn_set_signal(int signum)
