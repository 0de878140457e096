/* Plain-C restatement of gbnet's OR-NOR Gibbs sampler — TEST INFRASTRUCTURE ONLY.
 * See ornor_oracle.h for the rules on who may use it and how it is pinned.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/).  The node graph of the reference (heap objects + virtual calls)
 * is flattened into arrays, but every floating-point expression keeps the reference's
 * operand order, and the build uses -ffp-contract=off, so results agree with the
 * compiled reference to the last bit except where noted:
 *
 *  - XNode/TNode::get_children_loglikelihood sums over a std::set<HNode*>, i.e. in heap
 *    POINTER order (HParentNode.h:18, XNode.cpp:19-20); here the sum runs in ascending
 *    target order.  Differences are a few ulp of the sum.
 *  - lgamma/log/exp come from this host's libm rather than GSL's gsl_sf_lngamma.
 *
 * The random stream is NOT the reference's (MT19937 seeded from time(NULL),
 * ModelORNOR.cpp:24; GraphBase.cpp:8-9): draws come from counter-based Philox4x32-10
 * streams keyed by (seed, chain, sweep, node), identical to the CUDA path, or from
 * injected per-node arrays.
 */
#include "ornor_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

/* ------------------------------------------------------------------ keyed draws */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
 * SC'11).  Replaces GraphBase.cpp:8-9. */
void orn_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t) 0xD2511F53u * c0;
        uint64_t p1 = (uint64_t) 0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t) p1;
        uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t) p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* stream kinds (counter word 3, low byte) */
enum { KIND_X = 0, KIND_S = 1, KIND_TBETA = 2, KIND_TEXP = 3, KIND_INIT = 4 };

/* One 32-bit uniform per discrete node; four neighbouring X nodes share one Philox block
 * (S nodes: keyed_s below).
 * u = w / 2^32 in [0,1): the same resolution gsl_rng_uniform has for mt19937. */
static double keyed_u(uint64_t seed, unsigned int chain, unsigned int sweep, unsigned int node,
                      unsigned int kind)
{
    uint32_t ctr[4] = { node >> 2, sweep, chain, kind };
    uint32_t key[2] = { (uint32_t) seed, (uint32_t) (seed >> 32) };
    uint32_t out[4];
    orn_philox4x32_10(ctr, key, out);
    return (double) out[node & 3] * (1.0 / 4294967296.0);
}

/* S nodes: the uniform of the j-th parent edge (append_parent order) of target i; four consecutive
 * edges of one target share one Philox block: counter (i, sweep, chain, KIND_S | (j/4) << 8) */
static double keyed_s(uint64_t seed, unsigned int chain, unsigned int sweep, unsigned int target,
                      unsigned int j)
{
    uint32_t ctr[4] = { target, sweep, chain, KIND_S | ((j >> 2) << 8) };
    uint32_t key[2] = { (uint32_t) seed, (uint32_t) (seed >> 32) };
    uint32_t out[4];
    orn_philox4x32_10(ctr, key, out);
    return (double) out[j & 3] * (1.0 / 4294967296.0);
}

double orn_draw_u(uint64_t seed, unsigned int chain, unsigned int sweep, unsigned int node,
                  unsigned int kind)
{
    return keyed_u(seed, chain, sweep, node, kind);
}

/* sequential word stream private to one T update: blocks (tf, sweep, chain, KIND_TBETA | b<<8) */
typedef struct {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    unsigned int have, block;
} word_stream;

static void ws_init(word_stream *ws, uint64_t seed, unsigned int chain, unsigned int sweep,
                    unsigned int tf)
{
    ws->key[0] = (uint32_t) seed; ws->key[1] = (uint32_t) (seed >> 32);
    ws->ctr[0] = tf; ws->ctr[1] = sweep; ws->ctr[2] = chain; ws->ctr[3] = KIND_TBETA;
    ws->have = 0; ws->block = 0;
}

static double ws_uniform_pos(word_stream *ws)
{
    for (;;) {
        if (ws->have == 0) {
            ws->ctr[3] = KIND_TBETA | (ws->block << 8);
            orn_philox4x32_10(ws->ctr, ws->key, ws->buf);
            ws->block++;
            ws->have = 4;
        }
        uint32_t w = ws->buf[4 - ws->have];
        ws->have--;
        if (w != 0) return (double) w * (1.0 / 4294967296.0);
    }
}

/* polar Box-Muller, one deviate per call */
static double ws_gaussian(word_stream *ws)
{
    double x, y, r2;
    do {
        x = -1.0 + 2.0 * ws_uniform_pos(ws);
        y = -1.0 + 2.0 * ws_uniform_pos(ws);
        r2 = x * x + y * y;
    } while (r2 > 1.0 || r2 == 0.0);
    return y * sqrt(-2.0 * log(r2) / r2);
}

/* Marsaglia & Tsang (2000); a < 1 through Gamma(a+1) * U^(1/a) */
static double ws_gamma(word_stream *ws, double a)
{
    double boost = 1.0;
    if (a < 1.0) {
        double u = ws_uniform_pos(ws);
        boost = pow(u, 1.0 / a);
        a = a + 1.0;
    }
    double d = a - 1.0 / 3.0;
    double c = (1.0 / 3.0) / sqrt(d);
    double x, v, u;
    for (;;) {
        do {
            x = ws_gaussian(ws);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        u = ws_uniform_pos(ws);
        if (u < 1.0 - 0.0331 * (x * x) * (x * x)) break;
        if (log(u) < 0.5 * (x * x) + d * (1.0 - v + log(v))) break;
    }
    return boost * (d * v);
}

/* replaces gsl_ran_beta (Beta.cpp:56): X1/(X1+X2) with X1~Gamma(a), X2~Gamma(b) */
double orn_draw_beta(uint64_t seed, unsigned int chain, unsigned int sweep, unsigned int tf,
                     double a, double b)
{
    word_stream ws;
    ws_init(&ws, seed, chain, sweep, tf);
    double x1 = ws_gamma(&ws, a);
    double x2 = ws_gamma(&ws, b);
    return x1 / (x1 + x2);
}

/* replaces gsl_ran_exponential(rng, 1.0) (Beta.cpp:87): -log1p(-u) */
double orn_draw_exp(uint64_t seed, unsigned int chain, unsigned int sweep, unsigned int tf)
{
    uint32_t ctr[4] = { tf, sweep, chain, KIND_TEXP };
    uint32_t key[2] = { (uint32_t) seed, (uint32_t) (seed >> 32) };
    uint32_t out[4];
    orn_philox4x32_10(ctr, key, out);
    double u = (double) out[0] * (1.0 / 4294967296.0);
    return -log1p(-u);
}

/* ------------------------------------------------------------------ GSL pieces */
/* gsl_ran_beta_pdf (GSL 2.x randist/beta.c), called at Beta.cpp:50,82,83 */
double orn_beta_pdf(double x, double a, double b)
{
    if (x < 0 || x > 1) return 0;
    double gab = lgamma(a + b), ga = lgamma(a), gb = lgamma(b);
    if (x == 0.0 || x == 1.0) {
        if (a > 1.0 && b > 1.0) return 0.0;
        return exp(gab - ga - gb) * pow(x, a - 1) * pow(1 - x, b - 1);
    }
    return exp(gab - ga - gb + log(x) * (a - 1) + log1p(-x) * (b - 1));
}

/* gsl_rstat (GSL 2.x rstat/rstat.c): Welford mean / M2; used at RVNode.cpp:115-128 */
typedef struct { double mean, M2; size_t n; } rstat;
static void rstat_reset(rstat *w) { w->mean = 0.0; w->M2 = 0.0; w->n = 0; }
static void rstat_add(rstat *w, double x)
{
    double delta = x - w->mean;
    double n = (double) ++(w->n);
    double delta_n = delta / n;
    double term1 = delta * delta_n * (n - 1.0);
    w->mean += delta_n;
    w->M2 += term1;
}
static double rstat_mean(const rstat *w) { return w->mean; }
static double rstat_variance(const rstat *w) { return w->n > 1 ? w->M2 / ((double) w->n - 1.0) : 0.0; }
static double rstat_sd(const rstat *w) { return sqrt(rstat_variance(w)); }

/* ------------------------------------------------------------------ model */
typedef struct {
    unsigned int *x;      /* [nX]  0/1 */
    double *t;            /* [nX] */
    unsigned int *s;      /* [E] in SLOT order (by-target CSR) */
    rstat *st_x;          /* [nX*2] */
    rstat *st_t;          /* [nX] */
    rstat *st_s;          /* [E*3] slot order */
    unsigned int sweep;   /* sweeps drawn since construction (Philox counter word 1) */
    /* injection */
    double *inj_flat, *inj_beta, *inj_exp;
    size_t n_flat, n_beta, n_exp, i_flat, i_beta;
} chain_t;

struct orn_model {
    orn_params p;
    unsigned int nX, nH, E;
    unsigned int *tf_uid, *trg_uid;       /* sorted unique ids (GraphORNOR.cpp:31-42) */
    /* by-target CSR; parents in edge input order = append_parent order (GraphORNOR.cpp:237-263) */
    unsigned int *trg_ptr;                /* [nH+1] */
    unsigned int *par_tf;                 /* [E] slot -> TF index */
    unsigned int *par_edge;               /* [E] slot -> edge input index */
    unsigned int *slot_of_edge;           /* [E] */
    unsigned int *slot_trg;               /* [E] slot -> target index */
    unsigned int *mor_idx;                /* [E] slot -> mor+1 */
    /* by-TF CSR of UNIQUE children (HParentNode::children is a set), ascending target index */
    unsigned int *tf_ptr;                 /* [nX+1] */
    unsigned int *tf_child;
    unsigned int *n_h_child;              /* [nX] edge count (GraphORNOR.cpp:206-211) */
    unsigned int *y;                      /* [nH] observed state 0/1/2 (GraphORNOR.cpp:131-134) */
    double *zval;                         /* [nH] ZY or ZN value (GraphORNOR.cpp:257-260) */
    unsigned char *x_active_prior;        /* [nX] in active_tf_set (GraphORNOR.cpp:64-67) */
    double *t_a, *t_b, *t_mean;           /* [nX] (GraphORNOR.cpp:213-223) */
    double XPROB[2], XPROB_ACTIVE[2];     /* GraphORNOR.h:40-41 */
    double SPROB[3][3];                   /* GraphORNOR.cpp:25 */
    double YPROB[3];                      /* GraphORNOR.h:45, GraphORNOR.cpp:113-125 */
    double YNOISE[3][3];                  /* GraphORNOR.h:46-48 */
    unsigned int n_nodes;
    chain_t *chains;
};

static char g_error[512] = "";
const char *orn_last_error(void) { return g_error; }

static int g_signal = 0;                  /* ModelBase.cpp:3-33 */
void orn_set_signal(int s) { g_signal = s; }

static int cmp_uint(const void *a, const void *b)
{
    unsigned int x = *(const unsigned int *) a, y = *(const unsigned int *) b;
    return (x > y) - (x < y);
}

static unsigned int sorted_unique(const unsigned int *in, unsigned int n, unsigned int **out)
{
    unsigned int *tmp = (unsigned int *) malloc(sizeof(unsigned int) * (n ? n : 1));
    memcpy(tmp, in, sizeof(unsigned int) * n);
    qsort(tmp, n, sizeof(unsigned int), cmp_uint);
    unsigned int m = 0;
    for (unsigned int i = 0; i < n; i++)
        if (i == 0 || tmp[i] != tmp[i - 1]) tmp[m++] = tmp[i];
    *out = tmp;
    return m;
}

static long find_uid(const unsigned int *sorted, unsigned int n, unsigned int v)
{
    long lo = 0, hi = (long) n - 1;
    while (lo <= hi) {
        long mid = (lo + hi) / 2;
        if (sorted[mid] == v) return mid;
        if (sorted[mid] < v) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

/* restates GraphORNOR::build_structure (GraphORNOR.cpp:13-265), inference mode only */
orn_model *orn_model_create(unsigned int n_edge, const unsigned int *src, const unsigned int *trg,
                            const int *mor,
                            unsigned int n_evid, const unsigned int *evid_trg, const int *evid_val,
                            unsigned int n_active, const unsigned int *active_tf,
                            const orn_params *params)
{
    if (n_evid == 0) {
        snprintf(g_error, sizeof g_error,
                 "empty evidence selects the reference's simulation mode (GraphORNOR.cpp:28), which this oracle does not restate");
        return NULL;
    }
    for (unsigned int e = 0; e < n_edge; e++)
        if (mor[e] != -1 && mor[e] != 0 && mor[e] != 1) {      /* GraphORNOR.cpp:39-41 */
            snprintf(g_error, sizeof g_error, "MOR values can only be either -1, 0 or 1");
            return NULL;
        }
    for (unsigned int i = 0; i < n_evid; i++)
        if (evid_val[i] < -1 || evid_val[i] > 1) {
            snprintf(g_error, sizeof g_error, "evidence values can only be either -1, 0 or 1");
            return NULL;
        }

    orn_model *m = (orn_model *) calloc(1, sizeof(orn_model));
    m->p = *params;
    m->E = n_edge;
    m->nX = sorted_unique(src, n_edge, &m->tf_uid);
    m->nH = sorted_unique(trg, n_edge, &m->trg_uid);
    const unsigned int nX = m->nX, nH = m->nH, E = m->E;

    const double XPROB[2] = { 0.99, 0.01 }, XPROB_ACTIVE[2] = { 0.10, 0.90 };   /* GraphORNOR.h:40-41 */
    memcpy(m->XPROB, XPROB, sizeof XPROB);
    memcpy(m->XPROB_ACTIVE, XPROB_ACTIVE, sizeof XPROB_ACTIVE);
    const double YPROB[3] = { 0.05, 0.90, 0.05 };                               /* GraphORNOR.h:45 */
    memcpy(m->YPROB, YPROB, sizeof YPROB);
    const double YNOISE[3][3] = { { 0.945, 0.050, 0.005 },                      /* GraphORNOR.h:46-48 */
                                  { 0.050, 0.900, 0.050 },
                                  { 0.005, 0.050, 0.945 } };
    memcpy(m->YNOISE, YNOISE, sizeof YNOISE);
    memcpy(m->SPROB, params->sprior, sizeof(double) * 9);                       /* GraphORNOR.cpp:25 */

    m->x_active_prior = (unsigned char *) calloc(nX ? nX : 1, 1);
    for (unsigned int i = 0; i < n_active; i++) {
        long k = find_uid(m->tf_uid, nX, active_tf[i]);
        if (k >= 0) m->x_active_prior[k] = 1;
    }

    /* evidence (GraphORNOR.cpp:108-152) */
    if (params->comp_yprob) {
        unsigned int deg_counts[3] = { 0, 0, 0 };
        for (unsigned int i = 0; i < n_evid; i++) deg_counts[evid_val[i] + 1] += 1;
        m->YPROB[0] = (double) deg_counts[0] / nH;
        m->YPROB[2] = (double) deg_counts[2] / nH;
        m->YPROB[1] = (double) 1. - m->YPROB[0] - m->YPROB[2];
    }
    m->y = (unsigned int *) malloc(sizeof(unsigned int) * (nH ? nH : 1));
    for (unsigned int i = 0; i < nH; i++) m->y[i] = 1;                          /* 1 -> not DEG */
    for (unsigned int i = 0; i < n_evid; i++) {
        long h = find_uid(m->trg_uid, nH, evid_trg[i]);
        if (h >= 0) m->y[h] = (unsigned int) (evid_val[i] + 1);
    }

    /* Z nodes are constants at their prior mean (GraphORNOR.cpp:155-164, 257-260) */
    double z_mean = params->z_alpha / (params->z_alpha + params->z_beta);
    double z0_mean = params->z0_alpha / (params->z0_alpha + params->z0_beta);
    m->zval = (double *) malloc(sizeof(double) * (nH ? nH : 1));
    for (unsigned int i = 0; i < nH; i++) m->zval[i] = (m->y[i] != 1) ? z_mean : z0_mean;

    /* by-target CSR in edge input order */
    m->trg_ptr = (unsigned int *) calloc(nH + 1, sizeof(unsigned int));
    unsigned int *e_tf = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
    unsigned int *e_trg = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
    m->n_h_child = (unsigned int *) calloc(nX ? nX : 1, sizeof(unsigned int));
    for (unsigned int e = 0; e < E; e++) {
        e_tf[e] = (unsigned int) find_uid(m->tf_uid, nX, src[e]);
        e_trg[e] = (unsigned int) find_uid(m->trg_uid, nH, trg[e]);
        m->trg_ptr[e_trg[e] + 1]++;
        m->n_h_child[e_tf[e]]++;                                                /* GraphORNOR.cpp:206-211 */
    }
    for (unsigned int i = 0; i < nH; i++) m->trg_ptr[i + 1] += m->trg_ptr[i];
    m->par_tf = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
    m->par_edge = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
    m->slot_of_edge = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
    m->slot_trg = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
    m->mor_idx = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
    unsigned int *fill = (unsigned int *) calloc(nH ? nH : 1, sizeof(unsigned int));
    for (unsigned int e = 0; e < E; e++) {
        unsigned int i = e_trg[e];
        unsigned int q = m->trg_ptr[i] + fill[i]++;
        m->par_tf[q] = e_tf[e];
        m->par_edge[q] = e;
        m->slot_of_edge[e] = q;
        m->slot_trg[q] = i;
        m->mor_idx[q] = (unsigned int) (mor[e] + 1);
    }
    free(fill);

    /* by-TF CSR of unique children, ascending target index */
    m->tf_ptr = (unsigned int *) calloc(nX + 1, sizeof(unsigned int));
    m->tf_child = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
    {
        /* count unique (tf,target) pairs: walk targets ascending, mark last target seen per TF */
        long *last = (long *) malloc(sizeof(long) * (nX ? nX : 1));
        for (unsigned int k = 0; k < nX; k++) last[k] = -1;
        for (unsigned int i = 0; i < nH; i++)
            for (unsigned int q = m->trg_ptr[i]; q < m->trg_ptr[i + 1]; q++) {
                unsigned int k = m->par_tf[q];
                if (last[k] != (long) i) { last[k] = i; m->tf_ptr[k + 1]++; }
            }
        for (unsigned int k = 0; k < nX; k++) m->tf_ptr[k + 1] += m->tf_ptr[k];
        unsigned int *pos = (unsigned int *) calloc(nX ? nX : 1, sizeof(unsigned int));
        for (unsigned int k = 0; k < nX; k++) last[k] = -1;
        for (unsigned int i = 0; i < nH; i++)
            for (unsigned int q = m->trg_ptr[i]; q < m->trg_ptr[i + 1]; q++) {
                unsigned int k = m->par_tf[q];
                if (last[k] != (long) i) { last[k] = i; m->tf_child[m->tf_ptr[k] + pos[k]++] = i; }
            }
        free(pos); free(last);
    }
    free(e_tf); free(e_trg);

    /* T priors from the log out-degree (GraphORNOR.cpp:213-223); the ctor's t_alpha/t_beta
     * arguments are overwritten there, i.e. ignored */
    m->t_a = (double *) malloc(sizeof(double) * (nX ? nX : 1));
    m->t_b = (double *) malloc(sizeof(double) * (nX ? nX : 1));
    m->t_mean = (double *) malloc(sizeof(double) * (nX ? nX : 1));
    const unsigned int n_zt_groups = 10;
    for (unsigned int k = 0; k < nX; k++) {
        double lognchild = log(m->n_h_child[k]);
        double a = params->t_lmargin + fmax(n_zt_groups - lognchild - 1, 0.) * params->t_focus;
        double b = params->t_hmargin + lognchild * params->t_focus;
        m->t_a[k] = a; m->t_b[k] = b; m->t_mean[k] = a / (a + b);
    }

    m->n_nodes = nX + (params->const_params ? 0 : nX) + E;

    /* chains (ModelORNOR.cpp:23-28) */
    m->chains = (chain_t *) calloc(params->n_graphs ? params->n_graphs : 1, sizeof(chain_t));
    for (unsigned int c = 0; c < params->n_graphs; c++) {
        chain_t *ch = &m->chains[c];
        ch->x = (unsigned int *) malloc(sizeof(unsigned int) * (nX ? nX : 1));
        ch->t = (double *) malloc(sizeof(double) * (nX ? nX : 1));
        ch->s = (unsigned int *) malloc(sizeof(unsigned int) * (E ? E : 1));
        ch->st_x = (rstat *) calloc((size_t) nX * 2 + 1, sizeof(rstat));
        ch->st_t = (rstat *) calloc((size_t) nX + 1, sizeof(rstat));
        ch->st_s = (rstat *) calloc((size_t) E * 3 + 1, sizeof(rstat));
        for (unsigned int k = 0; k < nX; k++) {
            /* Multinomial ctor with rng draws from the prior (Multinomial.cpp:42-49: inside the
             * base-class ctor the blanket is just prob[v]); dice = u * cum[last] */
            const double *prob = m->x_active_prior[k] ? m->XPROB_ACTIVE : m->XPROB;
            double c0 = exp(log(prob[0]));
            double c1 = c0 + exp(log(prob[1]));
            double dice = 0. * (1 - keyed_u(params->seed, c, 0, k, KIND_INIT))
                        + c1 * keyed_u(params->seed, c, 0, k, KIND_INIT);
            ch->x[k] = dice < c0 ? 0u : 1u;
            ch->t[k] = m->t_mean[k];                            /* GraphORNOR.cpp:223-224 */
        }
        for (unsigned int q = 0; q < E; q++) ch->s[q] = m->mor_idx[q];          /* GraphORNOR.cpp:246 */
    }
    return m;
}

void orn_model_destroy(orn_model *m)
{
    if (!m) return;
    for (unsigned int c = 0; c < m->p.n_graphs; c++) {
        chain_t *ch = &m->chains[c];
        free(ch->x); free(ch->t); free(ch->s); free(ch->st_x); free(ch->st_t); free(ch->st_s);
        free(ch->inj_flat); free(ch->inj_beta); free(ch->inj_exp);
    }
    free(m->chains);
    free(m->tf_uid); free(m->trg_uid); free(m->trg_ptr); free(m->par_tf); free(m->par_edge);
    free(m->slot_of_edge); free(m->slot_trg); free(m->mor_idx); free(m->tf_ptr); free(m->tf_child);
    free(m->n_h_child); free(m->y); free(m->zval); free(m->x_active_prior);
    free(m->t_a); free(m->t_b); free(m->t_mean);
    free(m);
}

unsigned int orn_n_chains(const orn_model *m) { return m->p.n_graphs; }
unsigned int orn_n_nodes(const orn_model *m) { return m->n_nodes; }
unsigned int orn_n_tf(const orn_model *m) { return m->nX; }
unsigned int orn_n_targets(const orn_model *m) { return m->nH; }
unsigned int orn_n_edges(const orn_model *m) { return m->E; }
void orn_edge_slots(const orn_model *m, unsigned int *out) { memcpy(out, m->slot_of_edge, sizeof(unsigned int) * m->E); }
void orn_t_prior(const orn_model *m, unsigned int tf, double *a, double *b) { *a = m->t_a[tf]; *b = m->t_b[tf]; }

/* ------------------------------------------------------------------ likelihoods */
/* HNodeORNOR::get_model_likelihood (HNodeORNOR.cpp:40-106) for h = `value`, parents given
 * as arrays in append_parent order. */
static double model_likelihood(unsigned int value, unsigned int n_parent,
                               const double *tval, const unsigned int *xval, const unsigned int *sval,
                               double zvalue, const double *prob)
{
    double pr0, pr1, pr2, zcompl_pn;
    switch (value) {
    case 0:
        pr0 = 1.; zcompl_pn = 1.;
        for (unsigned int j = 0; j < n_parent; j++) {
            if (sval[j] == 0) {
                zcompl_pn *= (1. - zvalue);
                pr0 *= (1. - tval[j] * xval[j]) * zvalue;
            } else if (sval[j] == 2) {
                zcompl_pn *= (1. - zvalue);
            }
        }
        return (1. - pr0) * (1. - zcompl_pn) + zcompl_pn * prob[0];
    case 2:
        pr0 = 1.; pr2 = 1.; zcompl_pn = 1.;
        for (unsigned int j = 0; j < n_parent; j++) {
            if (sval[j] == 2) {
                zcompl_pn *= (1. - zvalue);
                pr2 *= (1. - tval[j] * xval[j]) * zvalue;
            } else if (sval[j] == 0) {
                zcompl_pn *= (1. - zvalue);
                pr0 *= (1. - tval[j] * xval[j]) * zvalue;
            }
        }
        return (pr0 - pr2 * pr0) * (1. - zcompl_pn) + zcompl_pn * prob[2];
    default:
        pr1 = 1.; zcompl_pn = 1.;
        for (unsigned int j = 0; j < n_parent; j++) {
            if (sval[j] != 1) {
                zcompl_pn *= (1. - zvalue);
                pr1 *= (1. - tval[j] * xval[j]) * zvalue;
            }
        }
        return pr1 * (1. - zcompl_pn) + zcompl_pn * prob[1];
    }
}

void orn_target_kat(unsigned int n_parent, const double *t, const unsigned int *x,
                    const unsigned int *s, double z, const double yprob[3],
                    double p_out[3], double l_out[3])
{
    const double YNOISE[3][3] = { { 0.945, 0.050, 0.005 }, { 0.050, 0.900, 0.050 }, { 0.005, 0.050, 0.945 } };
    for (unsigned int h = 0; h < 3; h++) p_out[h] = model_likelihood(h, n_parent, t, x, s, z, yprob);
    for (unsigned int y = 0; y < 3; y++) {
        double l = 0.;
        for (unsigned int h = 0; h < 3; h++) l += p_out[h] * YNOISE[h][y];
        l_out[y] = l;
    }
}

/* HNode::get_own_likelihood, latent branch (HNode.cpp:33-53) x YDataNode::get_own_likelihood
 * (YDataNode.cpp:26-29) for target i in chain ch */
static double target_likelihood(const orn_model *m, const chain_t *ch, unsigned int i)
{
    unsigned int q0 = m->trg_ptr[i], np = m->trg_ptr[i + 1] - q0;
    double tv[np ? np : 1];
    unsigned int xv[np ? np : 1];
    for (unsigned int j = 0; j < np; j++) {
        unsigned int k = m->par_tf[q0 + j];
        tv[j] = ch->t[k]; xv[j] = ch->x[k];
    }
    double likelihood = 0.;
    for (unsigned int h = 0; h < 3; h++)
        likelihood += model_likelihood(h, np, tv, xv, ch->s + q0, m->zval[i], m->YPROB)
                    * m->YNOISE[h][m->y[i]];
    return likelihood;
}

void orn_target_likelihoods(orn_model *m, unsigned int chain, double *out)
{
    for (unsigned int i = 0; i < m->nH; i++) out[i] = target_likelihood(m, &m->chains[chain], i);
}

/* XNode/TNode::get_children_loglikelihood (XNode.cpp:16-22, TNode.cpp:15-25) */
static double children_loglik(const orn_model *m, const chain_t *ch, unsigned int k)
{
    double loglik = 0.;
    for (unsigned int c = m->tf_ptr[k]; c < m->tf_ptr[k + 1]; c++)
        loglik += log(target_likelihood(m, ch, m->tf_child[c]));
    return loglik;
}

/* RVNode::get_blanket_loglikelihood (RVNode.cpp:94-106) of X_k at its current value */
static double x_blanket(const orn_model *m, const chain_t *ch, unsigned int k)
{
    const double *prob = m->x_active_prior[k] ? m->XPROB_ACTIVE : m->XPROB;
    double loglik = 0.;
    loglik += log(prob[ch->x[k]]);                              /* RVNode.cpp:82-85, Multinomial.cpp:51-54 */
    loglik += children_loglik(m, ch, k);
    return loglik;
}

/* S node in slot q: one child, its target (SNode.cpp:15-21) */
static double s_blanket(const orn_model *m, const chain_t *ch, unsigned int q)
{
    double loglik = 0.;
    loglik += log(m->SPROB[m->mor_idx[q]][ch->s[q]]);
    loglik += log(target_likelihood(m, ch, m->slot_trg[q]));
    return loglik;
}

/* T node (Beta.cpp:48-51 + TNode.cpp:15-25) */
static double t_blanket(const orn_model *m, const chain_t *ch, unsigned int k)
{
    double loglik = 0.;
    loglik += log(orn_beta_pdf(ch->t[k], m->t_a[k], m->t_b[k]));
    if (m->p.noise_listen_children) loglik += children_loglik(m, ch, k); else loglik += 0.;
    return loglik;
}

/* node index -> kind / local index, random_nodes order (GraphORNOR.cpp:57-73,213-233,235-264) */
static char node_kind(const orn_model *m, unsigned int node, unsigned int *local)
{
    if (node < m->nX) { *local = node; return 'X'; }
    node -= m->nX;
    if (!m->p.const_params) {
        if (node < m->nX) { *local = node; return 'T'; }
        node -= m->nX;
    }
    *local = node;          /* edge input index */
    return 'S';
}

double orn_blanket_loglik(orn_model *m, unsigned int chain, unsigned int node, double value)
{
    chain_t *ch = &m->chains[chain];
    unsigned int j; double ll;
    switch (node_kind(m, node, &j)) {
    case 'X': { unsigned int keep = ch->x[j]; ch->x[j] = (unsigned int) value; ll = x_blanket(m, ch, j); ch->x[j] = keep; return ll; }
    case 'T': { double keep = ch->t[j]; ch->t[j] = value; ll = t_blanket(m, ch, j); ch->t[j] = keep; return ll; }
    default:  { unsigned int q = m->slot_of_edge[j]; unsigned int keep = ch->s[q]; ch->s[q] = (unsigned int) value; ll = s_blanket(m, ch, q); ch->s[q] = keep; return ll; }
    }
}

void orn_all_conditionals(orn_model *m, unsigned int chain, double *out)
{
    for (unsigned int n = 0; n < m->n_nodes; n++) {
        unsigned int j;
        char kind = node_kind(m, n, &j);
        out[3 * n] = out[3 * n + 1] = out[3 * n + 2] = NAN;
        if (kind == 'X') { for (unsigned int v = 0; v < 2; v++) out[3 * n + v] = orn_blanket_loglik(m, chain, n, v); }
        else if (kind == 'S') { for (unsigned int v = 0; v < 3; v++) out[3 * n + v] = orn_blanket_loglik(m, chain, n, v); }
        else out[3 * n] = t_blanket(m, &m->chains[chain], j);
    }
}

double orn_joint_logprob(orn_model *m, unsigned int chain)
{
    const chain_t *ch = &m->chains[chain];
    double lp = 0.;
    for (unsigned int k = 0; k < m->nX; k++) {
        const double *prob = m->x_active_prior[k] ? m->XPROB_ACTIVE : m->XPROB;
        lp += log(prob[ch->x[k]]);
    }
    if (!m->p.const_params)
        for (unsigned int k = 0; k < m->nX; k++) lp += log(orn_beta_pdf(ch->t[k], m->t_a[k], m->t_b[k]));
    for (unsigned int e = 0; e < m->E; e++) {
        unsigned int q = m->slot_of_edge[e];
        lp += log(m->SPROB[m->mor_idx[q]][ch->s[q]]);
    }
    for (unsigned int i = 0; i < m->nH; i++) lp += log(target_likelihood(m, ch, i));
    return lp;
}

/* ------------------------------------------------------------------ state */
void orn_get_state(const orn_model *m, unsigned int chain, double *out)
{
    const chain_t *ch = &m->chains[chain];
    unsigned int o = 0;
    for (unsigned int k = 0; k < m->nX; k++) out[o++] = ch->x[k];
    if (!m->p.const_params) for (unsigned int k = 0; k < m->nX; k++) out[o++] = ch->t[k];
    for (unsigned int e = 0; e < m->E; e++) out[o++] = ch->s[m->slot_of_edge[e]];
}

void orn_set_state(orn_model *m, unsigned int chain, const double *in)
{
    chain_t *ch = &m->chains[chain];
    unsigned int o = 0;
    for (unsigned int k = 0; k < m->nX; k++) ch->x[k] = (unsigned int) in[o++];
    if (!m->p.const_params) for (unsigned int k = 0; k < m->nX; k++) ch->t[k] = in[o++];
    for (unsigned int e = 0; e < m->E; e++) ch->s[m->slot_of_edge[e]] = (unsigned int) in[o++];
}

/* ------------------------------------------------------------------ draws */
void orn_inject(orn_model *m, unsigned int chain,
                const double *flat_u, size_t n_flat, const double *beta_v, size_t n_beta,
                const double *exp_v, size_t n_exp)
{
    chain_t *ch = &m->chains[chain];
    free(ch->inj_flat); free(ch->inj_beta); free(ch->inj_exp);
    ch->inj_flat = (double *) malloc(sizeof(double) * (n_flat ? n_flat : 1));
    ch->inj_beta = (double *) malloc(sizeof(double) * (n_beta ? n_beta : 1));
    ch->inj_exp = (double *) malloc(sizeof(double) * (n_exp ? n_exp : 1));
    memcpy(ch->inj_flat, flat_u, sizeof(double) * n_flat);
    memcpy(ch->inj_beta, beta_v, sizeof(double) * n_beta);
    memcpy(ch->inj_exp, exp_v, sizeof(double) * n_exp);
    ch->n_flat = n_flat; ch->n_beta = n_beta; ch->n_exp = n_exp;
    ch->i_flat = 0; ch->i_beta = 0;
}

void orn_clear_injection(orn_model *m, unsigned int chain)
{
    chain_t *ch = &m->chains[chain];
    free(ch->inj_flat); free(ch->inj_beta); free(ch->inj_exp);
    ch->inj_flat = ch->inj_beta = ch->inj_exp = NULL;
    ch->n_flat = ch->n_beta = ch->n_exp = ch->i_flat = ch->i_beta = 0;
}

void orn_export_draws(const orn_model *m, unsigned int chain, unsigned int sweep0, unsigned int n,
                      double *flat_u, double *beta_v, double *exp_v)
{
    const unsigned int nX = m->nX, E = m->E;
    for (unsigned int i = 0; i < n; i++) {
        unsigned int sw = sweep0 + i;
        double *f = flat_u + (size_t) i * (nX + E);
        for (unsigned int k = 0; k < nX; k++) f[k] = keyed_u(m->p.seed, chain, sw, k, KIND_X);
        for (unsigned int e = 0; e < E; e++) {
            unsigned int q = m->slot_of_edge[e], i = m->slot_trg[q];
            f[nX + e] = keyed_s(m->p.seed, chain, sw, i, q - m->trg_ptr[i]);
        }
        for (unsigned int k = 0; k < nX; k++) {
            beta_v[(size_t) i * nX + k] = orn_draw_beta(m->p.seed, chain, sw, k, m->t_a[k], m->t_b[k]);
            exp_v[(size_t) i * nX + k] = orn_draw_exp(m->p.seed, chain, sw, k);
        }
    }
}

/* gsl_ran_flat(rng, 0, b) = 0*(1-u) + b*u (Multinomial.cpp:92) */
static double flat_draw(const orn_model *m, chain_t *ch, unsigned int chain, unsigned int node,
                        unsigned int kind, double b)
{
    double u;
    if (ch->inj_flat) {
        if (ch->i_flat >= ch->n_flat) { fprintf(stderr, "ornor_oracle: flat queue exhausted\n"); abort(); }
        u = ch->inj_flat[ch->i_flat++];
    } else if (kind == KIND_S) {
        unsigned int i = m->slot_trg[node];
        u = keyed_s(m->p.seed, chain, ch->sweep, i, node - m->trg_ptr[i]);
    } else {
        u = keyed_u(m->p.seed, chain, ch->sweep, node, kind);
    }
    return 0. * (1 - u) + b * u;
}

/* Multinomial::compute_outcome_likelihood + sample (Multinomial.cpp:56-108) */
static unsigned int multinomial_draw(unsigned int n_outcomes, const double *llik, const double *prob,
                                     double (*dice_fn)(void *, double), void *ctx, unsigned int current)
{
    double cum[3], sum_lik = 0.;
    for (unsigned int i = 0; i < n_outcomes; i++) {
        double lik = exp(llik[i]);
        sum_lik += lik;
        cum[i] = sum_lik;
    }
    if (!(sum_lik > 0.)) {                                       /* Multinomial.cpp:78-85 */
        sum_lik = 0.;
        for (unsigned int i = 0; i < n_outcomes; i++) { sum_lik += prob[i]; cum[i] = sum_lik; }
    }
    double dice = dice_fn(ctx, cum[n_outcomes - 1]);
    for (unsigned int i = 0; i < n_outcomes; i++)
        if (dice < cum[i]) return i;
    return current;                                              /* no branch taken: value kept */
}

typedef struct { const orn_model *m; chain_t *ch; unsigned int chain, node, kind; } dice_ctx;
static double dice_cb(void *p, double b)
{
    dice_ctx *d = (dice_ctx *) p;
    return flat_draw(d->m, d->ch, d->chain, d->node, d->kind, b);
}

/* Beta::metropolis_hastings_with_prior_proposal + Beta::sample (Beta.cpp:60-104) */
static void t_update(const orn_model *m, chain_t *ch, unsigned int chain, unsigned int k)
{
    double a = m->t_a[k], b = m->t_b[k], prior_mean = m->t_mean[k];
    double prev = ch->t[k];
    double prev_loglik = t_blanket(m, ch, k);

    double draw, expo = 0.; int have_expo = 0;
    if (ch->inj_beta) {
        if (ch->i_beta >= ch->n_beta) { fprintf(stderr, "ornor_oracle: beta queue exhausted\n"); abort(); }
        draw = ch->inj_beta[ch->i_beta];
        if (ch->i_beta < ch->n_exp) { expo = ch->inj_exp[ch->i_beta]; have_expo = 1; }
        ch->i_beta++;
    } else {
        draw = orn_draw_beta(m->p.seed, chain, ch->sweep, k, a, b);
    }
    double sample_dx = draw - prior_mean;
    ch->t[k] += sample_dx;

    double prop_loglik = t_blanket(m, ch, k);
    if (!(prop_loglik > -INFINITY)) {
        ch->t[k] = prev;
    } else if (prev_loglik > -INFINITY) {
        double q10 = orn_beta_pdf(prior_mean + sample_dx, a, b);
        double q01 = orn_beta_pdf(prior_mean - sample_dx, a, b);
        double log_q = log(q01 / q10);
        double logratio = prop_loglik - prev_loglik + log_q;
        int accept = logratio >= 0.;
        if (!accept) {
            if (!have_expo) expo = orn_draw_exp(m->p.seed, chain, ch->sweep, k);
            accept = logratio > -expo;
        }
        if (!accept) ch->t[k] = prev;
    }
    rstat_add(&ch->st_t[k], ch->t[k]);                          /* Beta.cpp:98-101 */
}

/* one sweep over random_nodes (GraphBase.cpp:22-27): X[sorted uid], T[sorted uid], S[input order] */
static void chain_sweep(const orn_model *m, unsigned int chain)
{
    chain_t *ch = &m->chains[chain];
    for (unsigned int k = 0; k < m->nX; k++) {
        const double *prob = m->x_active_prior[k] ? m->XPROB_ACTIVE : m->XPROB;
        unsigned int current = ch->x[k];
        double llik[2];
        for (unsigned int v = 0; v < 2; v++) { ch->x[k] = v; llik[v] = x_blanket(m, ch, k); }
        ch->x[k] = current;
        dice_ctx d = { m, ch, chain, k, KIND_X };
        ch->x[k] = multinomial_draw(2, llik, prob, dice_cb, &d, current);
        for (unsigned int v = 0; v < 2; v++) rstat_add(&ch->st_x[2 * k + v], ch->x[k] == v ? 1. : 0.);
    }
    if (!m->p.const_params)
        for (unsigned int k = 0; k < m->nX; k++) t_update(m, ch, chain, k);
    for (unsigned int e = 0; e < m->E; e++) {
        unsigned int q = m->slot_of_edge[e];
        unsigned int current = ch->s[q];
        double llik[3];
        for (unsigned int v = 0; v < 3; v++) { ch->s[q] = v; llik[v] = s_blanket(m, ch, q); }
        ch->s[q] = current;
        dice_ctx d = { m, ch, chain, q, KIND_S };
        ch->s[q] = multinomial_draw(3, llik, m->SPROB[m->mor_idx[q]], dice_cb, &d, current);
        for (unsigned int v = 0; v < 3; v++) rstat_add(&ch->st_s[3 * q + v], ch->s[q] == v ? 1. : 0.);
    }
    ch->sweep++;
}

void orn_sample_n(orn_model *m, unsigned int n)
{
    #pragma omp parallel for schedule(dynamic, 1)
    for (unsigned int c = 0; c < m->p.n_graphs; c++)
        for (unsigned int i = 0; i < n; i++) chain_sweep(m, c);
}

double orn_time_sample_n(orn_model *m, unsigned int n)
{
    double t0 = omp_get_wtime();
    orn_sample_n(m, n);
    return omp_get_wtime() - t0;
}

int orn_omp_max_threads(void) { return omp_get_max_threads(); }
void orn_omp_set_threads(int n) { omp_set_num_threads(n); }

/* ModelBase::sample (ModelBase.cpp:114-129), including the quirk that the batch just drawn
 * is never clipped: dN is updated AFTER sample_n(dN) ran */
void orn_sample(orn_model *m, unsigned int N, unsigned int dN)
{
    unsigned int n;
    double gr = INFINITY;
    for (n = 0; n < N && gr > 1.10 && g_signal == 0;
         dN = (dN < N - n ? dN : N - n), n += dN, gr = orn_max_gelman_rubin(m))
        orn_sample_n(m, dN);
    g_signal = 0;
}

void orn_burn_stats(orn_model *m)
{
    for (unsigned int c = 0; c < m->p.n_graphs; c++) {
        chain_t *ch = &m->chains[c];
        for (unsigned int i = 0; i < 2 * m->nX; i++) rstat_reset(&ch->st_x[i]);
        for (unsigned int i = 0; i < m->nX; i++) rstat_reset(&ch->st_t[i]);
        for (size_t i = 0; i < (size_t) 3 * m->E; i++) rstat_reset(&ch->st_s[i]);
    }
}

/* ------------------------------------------------------------------ summaries */
static const rstat *node_stat(const orn_model *m, unsigned int chain, unsigned int node, unsigned int stat)
{
    const chain_t *ch = &m->chains[chain];
    unsigned int j;
    switch (node_kind(m, node, &j)) {
    case 'X': return &ch->st_x[2 * j + stat];
    case 'T': return &ch->st_t[j];
    default:  return &ch->st_s[3 * (size_t) m->slot_of_edge[j] + stat];
    }
}

static unsigned int node_n_stats(const orn_model *m, unsigned int node)
{
    unsigned int j;
    switch (node_kind(m, node, &j)) { case 'X': return 2; case 'T': return 1; default: return 3; }
}

double orn_chain_length(const orn_model *m) { return (double) node_stat(m, 0, 0, 0)->n; }
double orn_node_mean(const orn_model *m, unsigned int chain, unsigned int node, unsigned int stat)
{ return rstat_mean(node_stat(m, chain, node, stat)); }
double orn_node_variance(const orn_model *m, unsigned int chain, unsigned int node, unsigned int stat)
{ return rstat_variance(node_stat(m, chain, node, stat)); }

/* ModelBase::get_gelman_rubin (ModelBase.cpp:47-104) */
void orn_gelman_rubin(const orn_model *m, double *out)
{
    double N = orn_chain_length(m);
    double M = m->p.n_graphs;
    if (!(N > 0)) { for (unsigned int i = 0; i < m->n_nodes; i++) out[i] = INFINITY; return; }
    for (unsigned int i = 0; i < m->n_nodes; i++) {
        double max_gr = 0.;
        unsigned int ns = node_n_stats(m, i);
        for (unsigned int j = 0; j < ns; j++) {
            rstat rB, rW; rstat_reset(&rB); rstat_reset(&rW);
            for (unsigned int k = 0; k < m->p.n_graphs; k++) {
                rstat_add(&rB, orn_node_mean(m, k, i, j));
                rstat_add(&rW, orn_node_variance(m, k, i, j));
            }
            double B = N * rstat_variance(&rB);
            double W = rstat_mean(&rW);
            double Vhat = W * (N - 1.) / N + B * (M + 1.) / (M * N);
            double gr;
            if (W > 0. && Vhat > 0.) gr = sqrt((N + 3.) / (N + 1.) * Vhat / W);
            else if (Vhat > 0.) gr = INFINITY;
            else gr = 1.;
            max_gr = fmax(gr, max_gr);
        }
        out[i] = max_gr;
    }
}

double orn_max_gelman_rubin(const orn_model *m)
{
    double *gr = (double *) malloc(sizeof(double) * (m->n_nodes ? m->n_nodes : 1));
    orn_gelman_rubin(m, gr);
    double mx = 0.;
    for (unsigned int i = 0; i < m->n_nodes; i++) mx = fmax(mx, gr[i]);
    free(gr);
    return mx;
}

/* ModelBase::get_posterior_means / get_posterior_sdevs (ModelBase.cpp:160-209): statistics of
 * the per-chain means over chains 1..M-1 — chain 0 is skipped by the reference */
static unsigned int posterior(const orn_model *m, const char *name, double *out, int want_sd)
{
    unsigned int o = 0;
    for (unsigned int i = 0; i < m->n_nodes; i++) {
        unsigned int j;
        if (node_kind(m, i, &j) != name[0] || name[1] != 0) continue;
        unsigned int ns = node_n_stats(m, i);
        for (unsigned int s = 0; s < ns; s++) {
            rstat r; rstat_reset(&r);
            for (unsigned int k = 1; k < m->p.n_graphs; k++) rstat_add(&r, orn_node_mean(m, k, i, s));
            if (out) out[o] = want_sd ? rstat_sd(&r) : rstat_mean(&r);
            o++;
        }
    }
    return o;
}

unsigned int orn_posterior_means(const orn_model *m, const char *name, double *out) { return posterior(m, name, out, 0); }
unsigned int orn_posterior_sdevs(const orn_model *m, const char *name, double *out) { return posterior(m, name, out, 1); }
