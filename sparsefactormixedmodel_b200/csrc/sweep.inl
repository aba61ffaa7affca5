// Stateful sweep: the loop body of fast_BSFG_sampler (R_BSFG/fast_BSFG_sampler.R:176-270) with the
// state resident in HBM and the data pre-rotated once (U'Y, Design_U'Y), so none of the reference's
// per-iteration n^2 p products remain.  Included by bsfg.cu.

#include <dlfcn.h>
#include <chrono>
#include <cstdlib>
#include <nccl.h>
#include <cuda_profiler_api.h>

namespace bsfg {

// ---- NCCL, resolved at run time so a single-GPU process never needs the library ---------------
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    void load() {
        if (handle) return;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) fail(BSFG_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define BSFG_NCCL_SYM(field, name)                                             \
        *(void**)(&field) = dlsym(handle, name);                               \
        if (!field) fail(BSFG_ENCCL, "symbol %s missing from libnccl", name)
        BSFG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        BSFG_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        BSFG_NCCL_SYM(AllReduce, "ncclAllReduce");
        BSFG_NCCL_SYM(AllGather, "ncclAllGather");
        BSFG_NCCL_SYM(Broadcast, "ncclBroadcast");
        BSFG_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        BSFG_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef BSFG_NCCL_SYM
    }
};
static NcclApi g_nccl;
#define BSFG_NCCL(expr)                                                                         \
    do {                                                                                        \
        ncclResult_t r__ = (expr);                                                              \
        if (r__ != ncclSuccess) fail(BSFG_ENCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(r__)); \
    } while (0)

enum Phase { PH_LAMBDA_PREP = 0, PH_LAMBDA_GRAM, PH_LAMBDA_CHOL, PH_MEANS, PH_H2, PH_FA, PH_F_GEMM, PH_F_COMM,
             PH_F_SOLVE, PH_PREC, PH_DELTA, PH_UPDATE_K };
static_assert(PH_UPDATE_K + 1 == BSFG_N_PHASES, "phase count");

}  // namespace bsfg

struct bsfg_ctx {
    bsfg::Device dev;
    bsfg_config cfg{};
    bsfg_priors pri{};
    std::vector<double> h2_prior_host;
    bool have_priors = false, have_consts = false, have_state = false;
    int n = 0, pl = 0, p_total = 0, p_off = 0, k = 0, kmax = 0, r = 0, b = 0, br = 0, H = 0;
    bool z1_identity = false;
    bool px_mode = false;         // parameter-expanded sampler: c->F holds F_px, c->Lam / c->Lt hold Lambda_px
    bsfg::DVec px, px_sq, px_isq, px_ipx, aux_px_rate;
    bsfg::DMat v_gpx;
    bool has_missing = false;     // NaN in Y: impute + re-rotate every iteration (fast_BSFG_sampler.R:180-184)
    bsfg::DMat Y0, Ycur, fit, Ft, Xt, v_zY, DtXt;
    int r2 = 0;                   // second random effect (Z_2 / W), 0: none
    bsfg::DMat Z2, Z2t, Uw, Uwt, A2inv, UtZ2t, DtZ2, DtZ2t, Z2tY, Z2tZ2;          // constants
    bsfg::DVec s1w, s2w, wp;
    bsfg::DMat W, Wt, UtYw, gW, tmpw, thw, Z2DUth, ZZW, A2W, FtZ2, Z2tF, FZ2W;    // state W + scratch
    bsfg::DMat v_zW, v_gW;
    bsfg::DVec aux_rw;
    int mode = 0;                 // 1 direct, 2 diag
    double resid_s1 = -1, resid_s2 = -1;
    long iter = 0;                // current_state$nrun
    // constants
    bsfg::DMat UtY, Yt, DtY, U1, UtX, U2, U2t, DU, DUt, U3, U3t, Z1, Z1t, M1, M2;
    bsfg::DVec yy, s, s1_2, s2_2, s1_3, s2_3, h2_prior;
    // state (factor-indexed storage has kmax columns / rows)
    bsfg::DMat Lt, Lam, Lmsg, LP, Plamt, Ptmp, F, Fa, Bfix, theta, thetat, Ea_in;
    bsfg::DVec F_h2, delta, tauh, rp, ep;
    bool theta_valid = false, derived_valid = false;
    bool bc_have_fa = false, bc_have_h2 = false;      // set_state with bcast_replicated: what rank 0 supplied
    // derived from F
    bsfg::DMat UtF, FtD, DtF;
    // scratch
    bsfg::LambdaScratch lsc;
    bsfg::GramOpts gram;
    bsfg::DMat tmp_brp, FDth, M1th, M2th, red1, YL, ZFa, T1, b3, vfa, FtF, FtY, G2;
    bsfg::DVec tau_e, Lp, Lfac, det, ll, red2, colsum_cnt, delta_shapes, delta_rates;
    int* d_ndraws = nullptr;
    int* d_map = nullptr;
    // injected-variate staging
    bsfg::DMat v_zL, v_zM, v_uh2, v_zFa, v_zF, v_gLP, v_gR, v_gE, v_gD, v_kadd;
    // aux
    bsfg::DMat aux_lp_rate;
    bsfg::DVec aux_rr, aux_re, aux_logps;
    bool aux_on = false;
    // device-side posterior (bsfg_posterior_begin / bsfg_sweep_collect)
    int post_cap = 0, post_n = 0;
    std::vector<int> post_spnum, post_k;
    bsfg::DMat tr_Lambda, tr_F, tr_Fa, tr_delta, tr_h2, tr_rp, tr_ep, tr_wp;     // one column per sample slot
    bsfg::DMat mean_theta, mean_W;                                                // running means (B, E_a through theta)
    double post_mean_count = 0;
    // multi-GPU
    ncclComm_t comm = nullptr;
    bsfg::DVec gstage;            // row-block exchange staging for the row-sharded replicated draws
    // side stream: the draws that depend on F only (F_h2, F_a) run beside the Lambda / theta chain, the delta chain beside the
    // precision statistics; its collectives use a communicator of their own (NCCL orders operations per communicator)
    bsfg::Device dev2;
    bool side_ready = false;
    ncclComm_t comm2 = nullptr;
    bsfg::DVec gstage2;
    cudaEvent_t ev_forkA = nullptr, ev_joinA = nullptr, ev_forkB = nullptr, ev_joinB = nullptr, ev_forkC = nullptr, ev_joinC = nullptr, ev_forkD = nullptr, ev_joinD = nullptr;
    // timing
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_phase;     // (N_PHASES + 1) per iteration slot, reused
    double last_total_ms = 0;
    long long last_launches = 0;
    double phase_ms[BSFG_N_PHASES] = {0};
    bool phase_timing = false;
    bool serial_iter = false;     // this iteration runs on the sweep stream alone (the phase-timed last iteration of a multi-iteration call)
};

namespace bsfg {

static void ctx_alloc_factor_storage(bsfg_ctx* c) {
    const int n = c->n, pl = c->pl, km = c->kmax, r = c->r, br = c->br;
    c->Lt.alloc(km, pl); c->Lam.alloc(pl, km); c->Lmsg.alloc(pl, km); c->LP.alloc(pl, km); c->Plamt.alloc(km, pl); c->Ptmp.alloc(pl, km);
    c->F.alloc(n, km); c->Fa.alloc(r, km); c->UtF.alloc(n, km); c->FtD.alloc(km, br); c->DtF.alloc(br, km);
    c->F_h2.alloc(km); c->delta.alloc(km); c->tauh.alloc(km); c->tau_e.alloc(km);
    c->px.alloc(km); c->px_sq.alloc(km); c->px_isq.alloc(km); c->px_ipx.alloc(km); c->aux_px_rate.alloc(km);
    c->YL.alloc(n, km); c->ZFa.alloc(n, km); c->T1.alloc(r, km); c->b3.alloc(r, km); c->vfa.alloc(r, km);
    c->FtF.alloc(km, km); c->FtY.alloc(km, pl); c->G2.alloc(km, pl); c->FDth.alloc(km, pl);
    c->Lp.alloc((long)km * (km + 1) / 2); c->Lfac.alloc((long)km * (km + 1) / 2);
    c->det.alloc(c->H); c->ll.alloc((long)km * c->H);
    // red1 = [LtL (km x km) | P1 (n x km) | P2 (br x km)] stacked by rows: one allreduce buffer
    c->red1.alloc(round_up(km, 2) + round_up(n, 2) + round_up(br, 2) + round_up(c->r2, 2), km);
    if (c->r2 > 0) { c->FtZ2.alloc(km, c->r2); c->Z2tF.alloc(c->r2, km); c->FZ2W.alloc(km, pl); }
    c->colsum_cnt.alloc(2 * (long)km);
    c->delta_shapes.alloc(km + 2); c->delta_rates.alloc(km + 2);
    c->aux_logps.alloc((long)km * c->H);
}

// view helpers: a DMat header over existing storage restricted to the first `cols` columns / `rows` rows
static DMat view(const DMat& m, long rows, long cols) {
    DMat v;
    v.p = m.p; v.rows = rows; v.cols = cols; v.ld = m.ld; v.owner = false;
    return v;
}
static DMat view_at(double* p, long rows, long cols, long ld) {
    DMat v;
    v.p = p; v.rows = rows; v.cols = cols; v.ld = ld; v.owner = false;
    return v;
}

static void refresh_derived(bsfg_ctx* c, bool fork_side = false);
// `side`: issue on the side stream with its own communicator (see bsfg_ctx::dev2)
static void allreduce(bsfg_ctx* c, double* buf, size_t count, bool side = false) {
    if (c->cfg.world <= 1) return;
    ncclComm_t comm = side ? c->comm2 : c->comm;
    if (!comm) fail(BSFG_ESTATE, "world > 1 but bsfg_comm_init was not called");
    BSFG_NCCL(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclSum, comm, side ? c->dev2.st : c->dev.st));
}

static void allgather_inplace(bsfg_ctx* c, double* buf, size_t chunk_doubles, bool side = false) {
    if (c->cfg.world <= 1) return;
    ncclComm_t comm = side ? c->comm2 : c->comm;
    if (!comm) fail(BSFG_ESTATE, "world > 1 but bsfg_comm_init was not called");
    BSFG_NCCL(g_nccl.AllGather(buf + (size_t)c->cfg.rank * chunk_doubles, buf, chunk_doubles, ncclDouble, comm,
                               side ? c->dev2.st : c->dev.st));
}

// Row sharding of the REPLICATED work (everything that depends on F only): with `world` ranks each computes rows
// [lo, lo+cnt) of an (rows x k) result -- rows of C = At'B are columns of At, a contiguous block -- and the blocks are
// exchanged with one all-gather, so every rank ends with bit-identical copies without redoing the other ranks' rows.
struct RowShard { long chunk, lo, cnt; };
static bool shard_replicated_enabled(int what /* 1 rows, 2 node product */) {
    // diagnostic switch: BSFG_SHARD_REPLICATED=0 none, 1 rows only, 2 node product only, 3 (default) both
    static const int mask = [] { const char* e = getenv("BSFG_SHARD_REPLICATED"); return e ? atoi(e) : 3; }();
    return (mask & what) != 0;
}
static RowShard row_shard(const bsfg_ctx* c, long rows) {
    RowShard s;
    if (c->cfg.world <= 1 || !shard_replicated_enabled(1)) { s.chunk = rows; s.lo = 0; s.cnt = rows; return s; }
    s.chunk = round_up(ceil_div(rows, c->cfg.world), 2);
    s.lo = (long)c->cfg.rank * s.chunk;
    s.cnt = std::max(0L, std::min(s.chunk, rows - s.lo));
    return s;
}
// C[lo:lo+cnt, :] (+)= alpha * At[:, lo:lo+cnt]' B
static void gemm_rows(bsfg_ctx* c, const RowShard& sh, double alpha, const DMat& At, const DMat& B, double beta, DMat& C, bool side = false) {
    if (sh.cnt <= 0) return;
    Device& d = side ? c->dev2 : c->dev;
    dgemm_tn(d.st, d.sms, (int)sh.cnt, (int)B.cols, (int)At.rows, alpha, At.p + sh.lo * At.ld, At.ld, B.p, B.ld, beta,
             C.p + sh.lo, C.ld, C.p + sh.lo, C.ld, d.ws.p, Device::WS_DOUBLES);
}
static void gather_rows(bsfg_ctx* c, const RowShard& sh, DMat& C, bool side = false) {
    if (c->cfg.world <= 1 || !shard_replicated_enabled(1)) return;
    Device& d = side ? c->dev2 : c->dev;
    DVec& stage = side ? c->gstage2 : c->gstage;
    const long cols = C.cols, per = sh.chunk * cols;
    if ((long)stage.n < per * c->cfg.world) stage.alloc(per * c->cfg.world);
    const int blocks = (int)std::min<long>((per + 255) / 256, (long)d.sms * 8);
    pack_rows_kernel<<<std::max(blocks, 1), 256, 0, d.st>>>(C.p, C.ld, sh.lo, sh.cnt, sh.chunk, cols, stage.p + (long)c->cfg.rank * per);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    allgather_inplace(c, stage.p, (size_t)per, side);
    const int blocks2 = (int)std::min<long>((per * c->cfg.world + 255) / 256, (long)d.sms * 8);
    unpack_rows_kernel<<<std::max(blocks2, 1), 256, 0, d.st>>>(stage.p, sh.chunk, cols, c->cfg.world, c->cfg.rank, C.rows, C.p, C.ld);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
}

// side stream, created on first use (BSFG_OVERLAP=0 keeps everything on the sweep stream)
static bool side_enabled(bsfg_ctx* c) {
    static const bool on = [] { const char* e = getenv("BSFG_OVERLAP"); return !e || atoi(e) != 0; }();
    if (!on) return false;
    if (c->cfg.world > 1 && !c->comm2) return false;
    if (!c->side_ready) {
        c->dev2.init();
        cudaFree(c->dev2.d_flag);
        c->dev2.d_flag = c->dev.d_flag;          // one not-positive-definite flag for the context
        // the staging buffer is sized before any concurrent use (a cudaMalloc in the middle of the sweep would serialise the streams)
        const long rows = std::max<long>(std::max(c->n, c->r), c->br);
        if (c->cfg.world > 1) c->gstage2.alloc(round_up(ceil_div(rows, c->cfg.world), 2) * c->kmax * c->cfg.world);
        for (cudaEvent_t* e : {&c->ev_forkA, &c->ev_joinA, &c->ev_forkB, &c->ev_joinB, &c->ev_forkC, &c->ev_joinC, &c->ev_forkD, &c->ev_joinD})
            BSFG_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        c->side_ready = true;
    }
    return true;
}

struct IterVariates {   // device-side sources for one iteration
    VarSrc zL, zM, uh2, zFa, zF, gLP, gR, gE, gD, zW, gW, zY, gPX;
};

static void phase_mark(bsfg_ctx* c, int idx) {
    if (c->phase_timing) BSFG_CUDA(cudaEventRecord(c->ev_phase[idx], c->dev.st));
}

static bool side_enabled(bsfg_ctx* c);
// fork_side: Design_U'F (with its all-gather and transpose) goes to the side stream while U'F is formed on the sweep stream;
// the caller waits for c->ev_joinD before it touches DtF / FtD
static void refresh_derived(bsfg_ctx* c, bool fork_side) {
    // UtF = U'F (t(FtU), cpp:107) and DtF = Design_U'F (+ F'Z_2): depend only on F -> row-sharded over the ranks
    Device& d = c->dev;
    DMat F = view(c->F, c->n, c->k), UtF = view(c->UtF, c->n, c->k), FtD = view(c->FtD, c->k, c->br);
    DMat DtF = view(c->DtF, c->br, c->k);
    const RowShard sn = row_shard(c, c->n), sb = row_shard(c, c->br);
    const bool side = fork_side && side_enabled(c);
    if (side) {
        BSFG_CUDA(cudaEventRecord(c->ev_forkD, d.st));
        BSFG_CUDA(cudaStreamWaitEvent(c->dev2.st, c->ev_forkD, 0));
    }
    Device& dd = side ? c->dev2 : d;
    gemm_rows(c, sb, 1.0, c->DU, F, 0.0, DtF, side);
    gather_rows(c, sb, DtF, side);
    dd.transpose(DtF, FtD);
    if (side) BSFG_CUDA(cudaEventRecord(c->ev_joinD, c->dev2.st));
    gemm_rows(c, sn, 1.0, c->U1, F, 0.0, UtF);
    gather_rows(c, sn, UtF);
    if (c->r2 > 0) {
        DMat FtZ2 = view(c->FtZ2, c->k, c->r2), Z2tF = view(c->Z2tF, c->r2, c->k);
        d.gemm(1.0, F, c->Z2, 0.0, FtZ2);          // F' Z_2
        d.transpose(FtZ2, Z2tF);
    }
    c->derived_valid = true;
}

// rotate the data: U'Y, (U'Y)', Design_U'Y, colSums(Y^2), Z_2'Y  (once at upload; every iteration with missing data)
static void rotate_data(bsfg_ctx* c, const DMat& Y) {
    Device& d = c->dev;
    d.gemm(1.0, c->U1, Y, 0.0, c->UtY);
    d.transpose(Y, c->Yt);                     // Y' (p x n): the K = p operand of Y_tilde Lmsg (cpp:149)
    d.gemm(1.0, c->DU, Y, 0.0, c->DtY);
    if (c->r2 > 0) d.gemm(1.0, c->Z2, Y, 0.0, c->Z2tY);
    colsumsq_kernel<<<ceil_div((long)c->pl * 32, 256), 256, 0, d.st>>>(Y.p, Y.ld, c->n, c->pl, c->yy.p);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
}

// step 0: Y[Y_missing] = (X B + F Lambda' + Z_1 E_a + Z_2 W + resids)[Y_missing]        R:180-184
static void impute_missing(bsfg_ctx* c, const VarSrc& zY) {
    Device& d = c->dev;
    const int n = c->n, pl = c->pl, k = c->k, r = c->r, b = c->b, r2 = c->r2;
    if (c->theta_valid) {
        d.gemm(1.0, c->DUt, c->theta, 0.0, c->fit);                       // Design_U theta = X B + Z_1 E_a
    } else {                                                              // first iteration after set_state
        if (!c->Ea_in.p) fail(BSFG_ESTATE, "Y has missing entries: bsfg_set_state must supply E_a");
        if (b > 0) d.gemm(1.0, c->Xt, c->Bfix, 0.0, c->fit);
        else c->fit.zero(d.st);
        if (c->z1_identity) {
            add_inplace_kernel<<<d.sms * 8, 256, 0, d.st>>>(c->fit.p, c->fit.ld, c->Ea_in.p, c->Ea_in.ld, n, pl);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        } else {
            d.gemm(1.0, c->Z1t, c->Ea_in, 1.0, c->fit);
        }
    }
    {
        DMat F = view(c->F, n, k), Ft = view(c->Ft, k, n), Lt = view(c->Lt, k, pl);
        d.transpose(F, Ft);
        d.gemm(1.0, Ft, Lt, 1.0, c->fit);                                 // + F Lambda'
    }
    if (r2 > 0) d.gemm(1.0, c->Z2t, c->W, 1.0, c->fit);                   // + Z_2 W
    impute_missing_kernel<<<pl, 256, 0, d.st>>>(c->Y0.p, c->Y0.ld, n, pl, c->fit.p, c->fit.ld, c->rp.p, zY, c->Ycur.p, c->Ycur.ld);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    rotate_data(c, c->Ycur);
    (void)r;
}

// ---- one Gibbs iteration (everything but update_k), steps numbered as SURVEY.md section 3.1 ----
static void gibbs_iteration(bsfg_ctx* c, const IterVariates& v, bool fixed_lambda = false, bool px = false) {
    Device& d = c->dev;
    const int n = c->n, pl = c->pl, k = c->k, r = c->r, b = c->b, br = c->br, H = c->H;
    if (!c->derived_valid) refresh_derived(c);
    DMat F = view(c->F, n, k), Fa = view(c->Fa, r, k), UtF = view(c->UtF, n, k), FtD = view(c->FtD, k, br);
    DMat Lt = view(c->Lt, k, pl), Lam = view(c->Lam, pl, k), Lmsg = view(c->Lmsg, pl, k), LP = view(c->LP, pl, k);
    DMat Plamt = view(c->Plamt, k, pl);

    // F_h2 | F and F_a | F, F_h2 (steps 4 and 5) depend on F only: with the side stream they start now and run beside the
    // Lambda / theta chain; the sweep stream picks their results up before the F draw.  (The px sweep keeps one stream: its
    // rescaled copies of F live in scratch the F step reuses.)
    const bool side = !px && !c->serial_iter && side_enabled(c);
    auto draw_h2_and_Fa = [&](Device& dd, bool on_side) {
        // 4. F_h2 | F (marginal over F_a)                         R:221 -> cpp:196-238
        //    px sweep: both draws see F_temp = F_px / sqrt(px_factor)  (fast_BSFG_sampler_px.R:232-238); YL / ZFa are free here
        DMat Fh = px ? view(c->YL, n, k) : view(c->F, n, k);
        DMat UtFh = px ? view(c->ZFa, n, k) : view(c->UtF, n, k);
        if (px) {
            px_vectors_kernel<<<ceil_div(k, 128), 128, 0, dd.st>>>(c->px.p, k, c->px_sq.p, c->px_isq.p, c->px_ipx.p);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
            dd.scale(F, nullptr, c->px_isq.p, Fh);
            dd.scale(UtF, nullptr, c->px_isq.p, UtFh);
        }
        h2s_core(dd, UtFh, Fh, c->s.p, H, c->h2_prior.p, v.uh2, c->F_h2.p, c->aux_on ? c->aux_logps.p : nullptr, c->det, c->ll);
        if (!on_side) phase_mark(c, 5);
        // 5. F_a | F, F_h2                                        R:226 -> cpp:293-339
        DMat T1 = view(c->T1, r, k), b3 = view(c->b3, r, k), vfa = view(c->vfa, r, k);
        const RowShard sr = row_shard(c, r);
        std::function<void(DMat&)> gather = [&](DMat& M) { gather_rows(c, sr, M, on_side); };
        fa_core(dd, Fh, c->z1_identity ? nullptr : &c->Z1, c->U3, c->U3t, c->s1_3.p, c->s2_3.p, c->F_h2.p, v.zFa, T1, b3, vfa, Fa,
                sr.lo, sr.cnt, (c->cfg.world > 1 && shard_replicated_enabled(1)) ? &gather : nullptr);
    };
    if (side) {
        BSFG_CUDA(cudaEventRecord(c->ev_forkA, d.st));
        BSFG_CUDA(cudaStreamWaitEvent(c->dev2.st, c->ev_forkA, 0));
        draw_h2_and_Fa(c->dev2, true);
        BSFG_CUDA(cudaEventRecord(c->ev_joinA, c->dev2.st));
    }
    // 0. missing phenotypes                                  R:180-184
    if (c->has_missing) impute_missing(c, v.zY);
    // 1. Lambda | F, B, precisions (marginal over E_a)      R:189-191 -> cpp:86-135
    phase_mark(c, 0);
    const int r2 = c->r2;
    if (!fixed_lambda) {
        if (r2 > 0)        // U'(Y - Z_2 W): R:189 (the X B part is fused into the weight kernel)
            dgemm_tn(d.st, d.sms, n, pl, r2, -1.0, c->UtZ2t.p, c->UtZ2t.ld, c->W.p, c->W.ld, 1.0, c->UtY.p, c->UtY.ld,
                     c->UtYw.p, c->UtYw.ld, d.ws.p, Device::WS_DOUBLES);
        lambda_core(d, c->lsc, c->gram, UtF, r2 > 0 ? c->UtYw : c->UtY, c->s.p, c->rp.p, c->ep.p, b > 0 ? &c->UtX : nullptr, b > 0 ? &c->Bfix : nullptr,
                    Plamt.p, Plamt.ld, 1, v.zL, Lt, c->phase_timing ? &c->ev_phase[1] : nullptr);
    } else if (c->phase_timing) {
        // fixed-Lambda sweep (fast_BSFG_sampler_fixedlambda.R:159-236): Lambda is a stored posterior draw, not resampled
        BSFG_CUDA(cudaEventRecord(c->ev_phase[1], d.st)); BSFG_CUDA(cudaEventRecord(c->ev_phase[2], d.st));
        BSFG_CUDA(cudaEventRecord(c->ev_phase[3], d.st));
    }
    {
        dim3 grid(ceil_div(k, 32), ceil_div(pl, 32)), block(32, 8);
        lt_to_lam_kernel<<<grid, block, 0, d.st>>>(Lt.p, Lt.ld, k, pl, c->rp.p, Lam.p, Lam.ld, Lmsg.p, Lmsg.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    // 2. [B; E_a] | F, Lambda in rotated coordinates theta      R:203-207 -> cpp:43-82
    //    Design_U'(Y - F Lambda') = DtY - (Design_U'F) Lambda'
    dgemm_tn(d.st, d.sms, br, pl, k, -1.0, FtD.p, FtD.ld, Lt.p, Lt.ld, 1.0, c->DtY.p, c->DtY.ld, c->tmp_brp.p,
             c->tmp_brp.ld, d.ws.p, Device::WS_DOUBLES);
    if (r2 > 0)        // ... - Design_U' Z_2 W  (R:203)
        dgemm_tn(d.st, d.sms, br, pl, r2, -1.0, c->DtZ2t.p, c->DtZ2t.ld, c->W.p, c->W.ld, 1.0, c->tmp_brp.p, c->tmp_brp.ld,
                 c->tmp_brp.p, c->tmp_brp.ld, d.ws.p, Device::WS_DOUBLES);
    if (fixed_lambda && b > 0)   // the fixed-Lambda driver also removes X B (old B) here: fast_BSFG_sampler_fixedlambda.R:180
        dgemm_tn(d.st, d.sms, br, pl, b, -1.0, c->DtXt.p, c->DtXt.ld, c->Bfix.p, c->Bfix.ld, 1.0, c->tmp_brp.p, c->tmp_brp.ld,
                 c->tmp_brp.p, c->tmp_brp.ld, d.ws.p, Device::WS_DOUBLES);
    {
        const bool fuse_b = b > 0 && b <= MEANS_BMAX;      // B = location_sample[1:b,] (R:206) reduced in the same pass
        // theta and theta' (the K = p operand of step 6) leave the same kernel: no transpose pass over the br x p array
        launch_means_theta(d.st, d.sms, d.means, c->tmp_brp.p, c->tmp_brp.ld, br, pl, c->s1_2.p, c->s2_2.p, c->rp.p, c->ep.p, v.zM,
                           c->theta.p, c->theta.ld, c->thetat.p, c->thetat.ld, c->U2t.p, c->U2t.ld, b,
                           fuse_b ? c->Bfix.p : nullptr, c->Bfix.ld);
        if (b > 0 && !fuse_b) {
            DMat U2Bt = view(c->U2t, br, b);               // first b columns of U2' = rows 0..b-1 of U2
            d.gemm(1.0, U2Bt, c->theta, 0.0, c->Bfix);
        }
    }
    c->theta_valid = true;
    // 3. W | B, E_a, F, Lambda: second sample_means_c call, with W_prec and the rand2 list    R:211-216
    //    Design_U2'(Y - X B - Z_1 E_a - F Lambda') = Uw'[Z_2'Y - (Z_2'Design_U) theta - (Z_2'F) Lambda']
    if (r2 > 0) {
        DMat FtZ2 = view(c->FtZ2, k, r2);
        d.gemm(1.0, c->DtZ2, c->theta, 0.0, c->Z2DUth);                                   // Z_2'Design_U theta (kept for step 8)
        dgemm_tn(d.st, d.sms, r2, pl, k, -1.0, FtZ2.p, FtZ2.ld, Lt.p, Lt.ld, 1.0, c->Z2tY.p, c->Z2tY.ld, c->gW.p, c->gW.ld,
                 d.ws.p, Device::WS_DOUBLES);                                             // Z_2'Y - Z_2'F Lambda'
        d.gemm(1.0, c->Uw, c->gW, 0.0, c->tmpw);
        d.gemm(-1.0, c->Uw, c->Z2DUth, 1.0, c->tmpw);
        launch_means_theta(d.st, d.sms, d.means, c->tmpw.p, c->tmpw.ld, r2, pl, c->s1w.p, c->s2w.p, c->rp.p, c->wp.p, v.zW, c->thw.p,
                           c->thw.ld, nullptr, 0, nullptr, 0, 0, nullptr, 0);
        d.gemm(1.0, c->Uwt, c->thw, 0.0, c->W);                                            // W = U theta_w (cpp:78)
        d.transpose(c->W, c->Wt);
    }
    phase_mark(c, 4);

    if (side) phase_mark(c, 5);             // steps 4 and 5 are on the side stream: their phases read ~0 on this one
    else draw_h2_and_Fa(d, false);
    phase_mark(c, 6);

    // 6. F | Lambda, B, E_a, F_a, F_h2                        R:230-232 -> cpp:138-163
    //    Y_tilde Lmsg = U (U'Y Lmsg) - Design_U (theta Lmsg);  the three partial products are the only
    //    cross-GPU traffic of the sweep (one allreduce of [Lambda'Lmsg | U'Y Lmsg | theta Lmsg]).
    const long o1 = round_up(c->kmax, 2), o2 = o1 + round_up(n, 2);
    DMat LtL = view_at(c->red1.p, k, k, c->red1.ld);
    DMat P1 = view_at(c->red1.p + o1, n, k, c->red1.ld);
    DMat P2 = view_at(c->red1.p + o2, br, k, c->red1.ld);
    d.gemm(1.0, Lam, Lmsg, 0.0, LtL);
    d.gemm(1.0, c->Yt, Lmsg, 0.0, P1);                      // Y Lmsg in original coordinates (U U' = I: no rotation needed here)
    d.gemm(1.0, c->thetat, Lmsg, 0.0, P2);
    const long o3 = o2 + round_up(br, 2);
    DMat P3 = view_at(c->red1.p + o3, r2, k, c->red1.ld);
    if (r2 > 0) d.gemm(1.0, c->Wt, Lmsg, 0.0, P3);                                         // W Lmsg
    phase_mark(c, 7);
    allreduce(c, c->red1.p, (size_t)c->red1.ld * (size_t)k);
    if (side) BSFG_CUDA(cudaStreamWaitEvent(d.st, c->ev_joinA, 0));        // F_h2 and F_a of this iteration are ready
    phase_mark(c, 8);
    {
        // rows of F are independent given the k x k factor: each rank forms and solves its own block of individuals,
        // then the blocks are all-gathered (bit-identical F on every rank)
        DMat YL = view(c->YL, n, k);
        const RowShard sn = row_shard(c, n);
        // the k x k factor of Lambda'Lmsg + diag(tau_e) (cpp:149) needs the reduced Lambda'Lmsg and F_h2 only: with the side
        // stream it is formed beside the Design_U product below
        auto factor_kxk = [&](Device& dd) {
            if (px) tau_e_px_kernel<<<ceil_div(k, 128), 128, 0, dd.st>>>(c->F_h2.p, c->px.p, k, c->tau_e.p);
            else tau_e_kernel<<<ceil_div(k, 128), 128, 0, dd.st>>>(c->F_h2.p, k, c->tau_e.p);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
            pack_lower_kernel<<<ceil_div((long)k * (k + 1) / 2, 256), 256, 0, dd.st>>>(LtL.p, LtL.ld, k, c->Lp.p);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
            VarSrc none = make_var(nullptr, 0, 0, 0, 0, 1);
            launch_chol_solve(dd, c->Lp.p, 0, k, 1, c->tau_e.p, 0, 1, nullptr, 0, none, nullptr, 0, c->Lfac.p);
        };
        if (side) {
            BSFG_CUDA(cudaEventRecord(c->ev_forkC, d.st));
            BSFG_CUDA(cudaStreamWaitEvent(c->dev2.st, c->ev_forkC, 0));
            factor_kxk(c->dev2);
            BSFG_CUDA(cudaEventRecord(c->ev_joinC, c->dev2.st));
        }
        if (sn.cnt > 0)                                                                  // YL = Y Lmsg - Design_U (theta Lmsg)
            dgemm_tn(d.st, d.sms, (int)sn.cnt, k, br, -1.0, c->DUt.p + sn.lo * c->DUt.ld, c->DUt.ld, P2.p, P2.ld, 1.0,
                     P1.p + sn.lo, P1.ld, YL.p + sn.lo, YL.ld, d.ws.p, Device::WS_DOUBLES);
        if (r2 > 0) gemm_rows(c, sn, -1.0, c->Z2t, P3, 1.0, YL);                         // ... - Z_2 W Lmsg  (R:230)
        // px sweep: F_a_px = F_a * sqrt(px), tau_e = 1 / (px (1 - F_h2))          fast_BSFG_sampler_px.R:249-250
        DMat Fa_use = px ? view(c->T1, r, k) : view(c->Fa, r, k);
        if (px) d.scale(Fa, nullptr, c->px_sq.p, Fa_use);
        const double* zfa_p = nullptr; long zfa_ld = 0;
        if (c->z1_identity) { zfa_p = Fa_use.p; zfa_ld = Fa_use.ld; }
        else {
            DMat ZFa = view(c->ZFa, n, k);
            if (px) d.gemm(1.0, c->Z1t, Fa_use, 0.0, ZFa);       // all rows: the px_factor draw needs Z_1 F_a_px in full
            else gemm_rows(c, sn, 1.0, c->Z1t, Fa_use, 0.0, ZFa);
            zfa_p = ZFa.p; zfa_ld = ZFa.ld;
        }
        if (side) BSFG_CUDA(cudaStreamWaitEvent(d.st, c->ev_joinC, 0));
        else factor_kxk(d);
        launch_factor_rows(d, c->Lfac.p, k, (int)sn.cnt, YL.p + sn.lo, YL.ld, zfa_p + sn.lo, zfa_ld, c->tau_e.p, v.zF,
                           F.p + sn.lo, F.ld, (int)sn.lo);
        gather_rows(c, sn, F);
        if (px) {
            // px_factor | F_px, F_a_px, F_h2                      fast_BSFG_sampler_px.R:260-263 (replicated on every rank)
            px_factor_kernel<<<k, 256, 0, d.st>>>(F.p, F.ld, zfa_p, zfa_ld, n, k, c->F_h2.p, c->pri.px_shape + n / 2.0, c->pri.px_rate,
                                                  v.gPX, c->px.p, c->aux_on ? c->aux_px_rate.p : nullptr);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
            px_vectors_kernel<<<ceil_div(k, 128), 128, 0, d.st>>>(c->px.p, k, c->px_sq.p, c->px_isq.p, c->px_ipx.p);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        }
    }
    phase_mark(c, 9);
    // px sweep: Lambda = Lambda_px * sqrt(px) feeds Lambda_prec, delta and update_k (:267); Lmsg is free from here on
    DMat LamS = px ? view(c->Lmsg, pl, k) : view(c->Lam, pl, k);
    if (px) d.scale(Lam, nullptr, c->px_sq.p, LamS);

    // 7. Lambda_prec | Lambda, tauh                           R:235-236
    if (!fixed_lambda) {
        const long total = (long)pl * k;
        int blocks = (int)std::min<long>((total + 255) / 256, (long)d.sms * 8);
        lambda_prec_kernel<<<blocks, 256, 0, d.st>>>(LamS.p, LamS.ld, pl, k, c->tauh.p, c->pri.Lambda_df, v.gLP, LP.p, LP.ld,
                                                     c->aux_on ? c->aux_lp_rate.p : nullptr);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    // 11-12. delta, tauh, Plam                                R:253-257 -> BSFG_functions.R:272-308
    //        needs Lambda and the new Lambda_prec only: with the side stream it runs beside the precision statistics below
    auto delta_chain = [&](Device& dd, bool on_side) {
        factor_reduce_kernel<<<k, 256, 0, dd.st>>>(LP.p, LP.ld, LamS.p, LamS.ld, pl, k, c->pri.epsilon, c->colsum_cnt.p,
                                                   c->colsum_cnt.p + c->kmax);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        allreduce(c, c->colsum_cnt.p, (size_t)2 * c->kmax, on_side);
        delta_recursion_kernel<<<1, 128, (4 * k + 2) * sizeof(double), dd.st>>>(c->delta.p, k, c->colsum_cnt.p, (double)c->p_total,
                                                                    c->pri.delta_1_shape, c->pri.delta_1_rate,
                                                                    c->pri.delta_2_shape, c->pri.delta_2_rate, v.gD, c->tauh.p,
                                                                    c->delta_shapes.p, c->delta_rates.p, c->d_ndraws);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        dim3 grid(ceil_div(pl, 32), ceil_div(k, 32)), block(32, 8);
        plam_kernel<<<grid, block, 0, dd.st>>>(LP.p, LP.ld, pl, k, c->tauh.p, Plamt.p, Plamt.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        if (px) dd.scale(Plamt, c->px_ipx.p, nullptr, Plamt);      // Plam = Plam / px_factor (:290); rows of Plamt are factors
    };
    if (side && !fixed_lambda) {
        BSFG_CUDA(cudaEventRecord(c->ev_forkB, d.st));
        BSFG_CUDA(cudaStreamWaitEvent(c->dev2.st, c->ev_forkB, 0));
        delta_chain(c->dev2, true);
        BSFG_CUDA(cudaEventRecord(c->ev_joinB, c->dev2.st));
    }
    // 8-9. resid_Y_prec, E_a_prec from sufficient statistics of the NEW F   R:239-245
    refresh_derived(c, /*fork_side=*/side);
    {
        DMat FtF = view(c->FtF, k, k), FtY = view(c->FtY, k, pl), G2 = view(c->G2, k, pl);
        d.gemm(1.0, F, F, 0.0, FtF);
        d.gemm(1.0, UtF, c->UtY, 0.0, FtY);                // F'Y = (U'F)'(U'Y)
        d.gemm(1.0, FtF, Lt, 0.0, G2);                     // F'F lambda_j
        if (side) BSFG_CUDA(cudaStreamWaitEvent(d.st, c->ev_joinD, 0));     // Design_U'F from the side stream
        DMat DtF = view(c->DtF, br, k), FDth = view(c->FDth, k, pl);
        d.gemm(1.0, DtF, c->theta, 0.0, FDth);             // (F' Design_U) theta_j: k values per trait
        if (c->mode == 1) {
            d.gemm(1.0, c->M1, c->theta, 0.0, c->M1th);
            d.gemm(1.0, c->M2, c->theta, 0.0, c->M2th);
        }
        TraitReduceArgs a;
        memset(&a, 0, sizeof a);
        if (r2 > 0) {
            DMat Z2tF = view(c->Z2tF, r2, k), FZ2W = view(c->FZ2W, k, pl);
            d.gemm(1.0, Z2tF, c->W, 0.0, FZ2W);              // F'Z_2 W
            d.gemm(1.0, c->Z2tZ2, c->W, 0.0, c->ZZW);        // Z_2'Z_2 W
            d.gemm(1.0, c->A2inv, c->W, 0.0, c->A2W);        // A_2_inv W  (R:249)
            a.W = c->W.p; a.ldw = c->W.ld; a.r2 = r2;
            a.Z2tY = c->Z2tY.p; a.ldzy = c->Z2tY.ld; a.Z2DUth = c->Z2DUth.p; a.ldzd = c->Z2DUth.ld;
            a.ZZW = c->ZZW.p; a.ldzz = c->ZZW.ld; a.A2W = c->A2W.p; a.lda2 = c->A2W.ld; a.FZ2W = FZ2W.p; a.ldfz = FZ2W.ld;
            a.wp = c->wp.p; a.rw_shape = c->pri.W_prec_shape; a.rw_rate = c->pri.W_prec_rate; a.r2_half = r2 / 2.0;
            a.rw_rate_out = c->aux_on ? c->aux_rw.p : nullptr;
        }
        a.theta = c->theta.p; a.ldth = c->theta.ld; a.DtY = c->DtY.p; a.lddy = c->DtY.ld;
        a.FDth = FDth.p; a.ldfd = FDth.ld;
        a.M1th = c->mode == 1 ? c->M1th.p : nullptr; a.M2th = c->mode == 1 ? c->M2th.p : nullptr; a.ldm = c->M1th.ld;
        a.s1 = c->s1_2.p; a.s2 = c->s2_2.p;
        a.Lt = Lt.p; a.ldlt = Lt.ld; a.FtY = FtY.p; a.ldfy = FtY.ld; a.G2 = G2.p; a.ldg2 = G2.ld;
        a.Bfix = c->Bfix.p; a.ldb = c->Bfix.ld; a.yy = c->yy.p;
        a.br = br; a.k = k; a.b = b; a.p = pl;
        a.n_half = n / 2.0; a.r_half = r / 2.0;
        a.rr_shape = c->pri.resid_Y_prec_shape; a.rr_rate = c->pri.resid_Y_prec_rate;
        a.re_shape = c->pri.E_a_prec_shape; a.re_rate = c->pri.E_a_prec_rate; a.bX = c->pri.b_X_prec;
        a.rp = c->rp.p; a.ep = c->ep.p;
        a.rr_rate_out = c->aux_on ? c->aux_rr.p : nullptr; a.re_rate_out = c->aux_on ? c->aux_re.p : nullptr;
        trait_reduce_kernel<<<ceil_div((long)pl * 32, 256), 256, 0, d.st>>>(a, v.gR, v.gE, v.gW);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    phase_mark(c, 10);

    if (fixed_lambda) { phase_mark(c, 11); return; }
    if (side) BSFG_CUDA(cudaStreamWaitEvent(d.st, c->ev_joinB, 0));       // delta, tauh, Plam of this iteration are in place
    else delta_chain(d, false);
    phase_mark(c, 11);
}

// ---- update_k (BSFG_functions.R:311-372) on a context: the shared implementation is update_k_apply (stateless.inl) ----
static void update_k_step(bsfg_ctx* c, long i, double uu, const KAddVariates* inj) {
    KRefs s;
    s.d = &c->dev; s.pri = &c->pri; s.seed = c->cfg.seed;
    s.n = c->n; s.pl = c->pl; s.r = c->r; s.kmax = c->kmax; s.p_total = c->p_total; s.p_off = c->p_off; s.k = &c->k;
    s.z1_identity = c->z1_identity; s.px_mode = c->px_mode;
    s.Z1 = c->z1_identity ? nullptr : c->Z1.p; s.ldz1 = c->Z1.ld;
    s.Lam = &c->Lam; s.LP = &c->LP; s.Lt = &c->Lt; s.Plamt = &c->Plamt; s.F = &c->F; s.Fa = &c->Fa;
    s.F_h2 = c->F_h2.p; s.delta = c->delta.p; s.tauh = c->tauh.p; s.px = c->px.p;
    s.cnt = c->colsum_cnt.p + c->kmax;        // reduced over ranks in step 11 of gibbs_iteration
    s.d_map = c->d_map;
    if (update_k_apply(s, i, uu, inj)) c->derived_valid = false;
}

static void require_ready(bsfg_ctx* c) {
    if (!c) fail(BSFG_EARG, "ctx is NULL");
    if (!c->have_priors) fail(BSFG_ESTATE, "bsfg_set_priors has not been called");
    if (!c->have_consts) fail(BSFG_ESTATE, "bsfg_upload_constants has not been called");
    if (!c->have_state) fail(BSFG_ESTATE, "bsfg_set_state has not been called");
    if (c->k < 2) fail(BSFG_ESHAPE, "k >= 2 required");
}

static void finish_timing(bsfg_ctx* c, long long launches_before, int n_iter_timed_phases) {
    Device& d = c->dev;
    BSFG_CUDA(cudaEventRecord(c->ev_end, d.st));
    BSFG_CUDA(cudaEventSynchronize(c->ev_end));
    float ms = 0;
    BSFG_CUDA(cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end));
    c->last_total_ms = ms;
    c->last_launches = g_launch_counter - launches_before;
    if (c->phase_timing && n_iter_timed_phases > 0) {
        // marks 0..11 bracket the 11 phases PH_LAMBDA_PREP .. PH_DELTA in order; PH_LAMBDA_GRAM is exactly the
        // weighted-Gram dgemm_tn launch (the kernel bench.py reports the roofline for).
        for (int q = 0; q < 11; q++) {
            float t = 0;
            BSFG_CUDA(cudaEventElapsedTime(&t, c->ev_phase[q], c->ev_phase[q + 1]));
            c->phase_ms[q] = t;
        }
    }
    check_notpd(d, "bsfg_sweep");
}

}  // namespace bsfg

extern "C" long long bsfg_launch_count(void) { return bsfg::g_launch_counter; }
extern "C" int bsfg_k_limit(void) { return bsfg::BSFG_K_LIMIT; }

extern "C" int bsfg_create(const bsfg_config* cfg, bsfg_ctx** out) {
    BSFG_API_BEGIN
    REQUIRE_PTR(cfg); REQUIRE_PTR(out);
    if (cfg->n <= 0 || cfg->p_local <= 0 || cfg->p_total < cfg->p_local || cfg->k < 2 || cfg->k_max < cfg->k || cfg->r <= 0 ||
        cfg->b < 0 || cfg->r2 < 0 || cfg->h2_divisions <= 0 || cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world ||
        cfg->p_offset < 0 || cfg->p_offset + cfg->p_local > cfg->p_total)
        fail(BSFG_ESHAPE, "bsfg_create: inconsistent sizes");
    if (cfg->z1_is_identity && cfg->r != cfg->n) fail(BSFG_ESHAPE, "z1_is_identity requires r == n");
    kt_for(cfg->k_max);
    GramOpts gram = resolve_gram(cfg->gram_mode, cfg->gram_terms < 0 ? 0 : cfg->gram_terms, cfg->gram_step);
    bsfg_ctx* c = new bsfg_ctx();
    try {
        c->dev.init();
        c->cfg = *cfg;
        c->gram = gram;
        c->gram.auto_terms = (cfg->gram_terms < 0 && gram.mode == GRAM_EXPSUM);   // opt-in: see bsfg_config
        c->gram.terms_max = gram.terms;
        c->n = cfg->n; c->pl = cfg->p_local; c->p_total = cfg->p_total; c->p_off = cfg->p_offset;
        c->k = cfg->k; c->kmax = cfg->k_max; c->r = cfg->r; c->b = cfg->b; c->br = cfg->b + cfg->r; c->H = cfg->h2_divisions;
        c->z1_identity = cfg->z1_is_identity != 0;
        c->r2 = cfg->r2;
        ctx_alloc_factor_storage(c);
        const int n = c->n, pl = c->pl, br = c->br, b = c->b;
        c->Bfix.alloc(std::max(b, 1), pl); c->theta.alloc(br, pl); c->thetat.alloc(pl, br);
        c->tmp_brp.alloc(br, pl);
        c->rp.alloc(pl); c->ep.alloc(pl); c->yy.alloc(pl);
        if (c->r2 > 0) {
            const int r2 = c->r2;
            c->wp.alloc(pl); c->aux_rw.alloc(pl);
            c->W.alloc(r2, pl); c->Wt.alloc(pl, r2); c->UtYw.alloc(n, pl); c->gW.alloc(r2, pl); c->tmpw.alloc(r2, pl);
            c->thw.alloc(r2, pl); c->Z2DUth.alloc(r2, pl); c->ZZW.alloc(r2, pl); c->A2W.alloc(r2, pl);
        }
        c->aux_lp_rate.alloc(pl, c->kmax); c->aux_rr.alloc(pl); c->aux_re.alloc(pl);
        BSFG_CUDA(cudaMalloc(&c->d_ndraws, sizeof(int)));
        BSFG_CUDA(cudaMalloc(&c->d_map, sizeof(int) * c->kmax));
        BSFG_CUDA(cudaEventCreate(&c->ev_begin));
        BSFG_CUDA(cudaEventCreate(&c->ev_end));
        c->ev_phase.resize(BSFG_N_PHASES + 1);
        for (auto& e : c->ev_phase) BSFG_CUDA(cudaEventCreate(&e));
        (void)n;
    } catch (...) {
        delete c;
        throw;
    }
    *out = c;
    BSFG_API_END
}

extern "C" int bsfg_destroy(bsfg_ctx* c) {
    BSFG_API_BEGIN
    if (!c) return BSFG_OK;
    cudaStreamSynchronize(c->dev.st);
    if (c->side_ready) {
        cudaStreamSynchronize(c->dev2.st);
        for (cudaEvent_t e : {c->ev_forkA, c->ev_joinA, c->ev_forkB, c->ev_joinB, c->ev_forkC, c->ev_joinC, c->ev_forkD, c->ev_joinD})
            if (e) cudaEventDestroy(e);
        if (c->dev2.means.Bpart) cudaFree(c->dev2.means.Bpart);
        if (c->dev2.means.counter) cudaFree(c->dev2.means.counter);
        if (c->dev2.st) cudaStreamDestroy(c->dev2.st);
    }
    if (c->comm2) g_nccl.CommDestroy(c->comm2);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->dev.means.Bpart) cudaFree(c->dev.means.Bpart);
    if (c->dev.means.counter) cudaFree(c->dev.means.counter);
    if (c->d_ndraws) cudaFree(c->d_ndraws);
    if (c->d_map) cudaFree(c->d_map);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    for (auto& e : c->ev_phase) cudaEventDestroy(e);
    if (c->dev.d_flag) cudaFree(c->dev.d_flag);
    if (c->dev.st) cudaStreamDestroy(c->dev.st);
    delete c;
    BSFG_API_END
}

extern "C" int bsfg_set_priors(bsfg_ctx* c, const bsfg_priors* pr) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c); REQUIRE_PTR(pr); REQUIRE_PTR(pr->h2_priors_factors);
    c->pri = *pr;
    c->h2_prior_host.assign(pr->h2_priors_factors, pr->h2_priors_factors + c->H);
    c->pri.h2_priors_factors = c->h2_prior_host.data();
    c->h2_prior.alloc(c->H);
    c->h2_prior.upload(c->h2_prior_host.data(), c->dev.st);
    c->dev.sync();
    c->have_priors = true;
    BSFG_API_END
}

extern "C" int bsfg_upload_constants(bsfg_ctx* c, const bsfg_constants* k) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c); REQUIRE_PTR(k);
    if (!c->have_priors) fail(BSFG_ESTATE, "call bsfg_set_priors before bsfg_upload_constants (b_X_prec is needed)");
    REQUIRE_PTR(k->Y); REQUIRE_PTR(k->U1); REQUIRE_PTR(k->s); REQUIRE_PTR(k->U2); REQUIRE_PTR(k->s1_2); REQUIRE_PTR(k->s2_2);
    REQUIRE_PTR(k->Design_U); REQUIRE_PTR(k->U3); REQUIRE_PTR(k->s1_3); REQUIRE_PTR(k->s2_3);
    Device& d = c->dev;
    const int n = c->n, pl = c->pl, r = c->r, b = c->b, br = c->br;
    if (b > 0) REQUIRE_PTR(k->X);
    if (!c->z1_identity) REQUIRE_PTR(k->Z_1);
    c->U1.alloc(n, n); c->U2.alloc(br, br); c->U2t.alloc(br, br); c->DU.alloc(n, br); c->DUt.alloc(br, n);
    c->U3.alloc(r, r); c->U3t.alloc(r, r);
    c->s.alloc(n); c->s1_2.alloc(br); c->s2_2.alloc(br); c->s1_3.alloc(r); c->s2_3.alloc(r);
    c->U1.upload(k->U1, d.st); c->U2.upload(k->U2, d.st); c->DU.upload(k->Design_U, d.st); c->U3.upload(k->U3, d.st);
    host_minmax(k->s, n, &c->gram.s_min, &c->gram.s_max);
    c->s.upload(k->s, d.st); c->s1_2.upload(k->s1_2, d.st); c->s2_2.upload(k->s2_2, d.st);
    c->s1_3.upload(k->s1_3, d.st); c->s2_3.upload(k->s2_3, d.st);
    d.transpose(c->U2, c->U2t); d.transpose(c->DU, c->DUt); d.transpose(c->U3, c->U3t);
    if (!c->z1_identity) {
        c->Z1.alloc(n, r); c->Z1t.alloc(r, n);
        c->Z1.upload(k->Z_1, d.st);
        d.transpose(c->Z1, c->Z1t);
    }
    if (c->r2 > 0) {
        const int r2 = c->r2;
        REQUIRE_PTR(k->Z_2); REQUIRE_PTR(k->U4); REQUIRE_PTR(k->s1_4); REQUIRE_PTR(k->s2_4);
        c->Z2.alloc(n, r2); c->Z2t.alloc(r2, n); c->Uw.alloc(r2, r2); c->Uwt.alloc(r2, r2); c->A2inv.alloc(r2, r2);
        c->s1w.alloc(r2); c->s2w.alloc(r2);
        c->Z2.upload(k->Z_2, d.st); c->Uw.upload(k->U4, d.st); c->s1w.upload(k->s1_4, d.st); c->s2w.upload(k->s2_4, d.st);
        if (k->A_2_inv) c->A2inv.upload(k->A_2_inv, d.st);
        else {
            std::vector<double> eye((size_t)r2 * r2, 0.0);
            for (int i = 0; i < r2; i++) eye[(size_t)i * r2 + i] = 1.0;      // A_2_inv = diag(1, r2), init:264
            c->A2inv.upload(eye.data(), d.st);
            d.sync();
        }
        d.transpose(c->Z2, c->Z2t); d.transpose(c->Uw, c->Uwt);
        DMat UtZ2(n, r2);
        c->UtZ2t.alloc(r2, n); c->DtZ2.alloc(br, r2); c->DtZ2t.alloc(r2, br); c->Z2tZ2.alloc(r2, r2);
        d.gemm(1.0, c->U1, c->Z2, 0.0, UtZ2);      d.transpose(UtZ2, c->UtZ2t);          // U'Z_2
        d.gemm(1.0, c->DU, c->Z2, 0.0, c->DtZ2);   d.transpose(c->DtZ2, c->DtZ2t);       // Design_U'Z_2
        d.gemm(1.0, c->Z2, c->Z2, 0.0, c->Z2tZ2);                                          // Z_2'Z_2
        d.sync();
    }
    {   // rotate the data once: U'Y, (U'Y)', Design_U'Y, colSums(Y^2), U'X
        c->has_missing = false;
        const size_t total = (size_t)n * (size_t)pl;
        for (size_t e = 0; e < total; e++)
            if (k->Y[e] != k->Y[e]) { c->has_missing = true; break; }       // is.na(Y), fast_BSFG_sampler_init.R:69
        c->UtY.alloc(n, pl); c->Yt.alloc(pl, n); c->DtY.alloc(br, pl);
        if (c->r2 > 0) c->Z2tY.alloc(c->r2, pl);
        if (c->has_missing) {
            // rotated data are rebuilt after every imputation; keep Y itself (NaN = missing) on the device
            c->Y0.alloc(n, pl); c->Ycur.alloc(n, pl); c->fit.alloc(n, pl); c->Ft.alloc(c->kmax, n);
            c->Y0.upload(k->Y, d.st);
            if (b > 0) {
                DMat X(n, b);
                X.upload(k->X, d.st);
                c->Xt.alloc(b, n);
                d.transpose(X, c->Xt);
                d.sync();
            }
        } else {
            DMat Y(n, pl);
            Y.upload(k->Y, d.st);
            rotate_data(c, Y);
            d.sync();
        }
        if (b > 0) {
            DMat X(n, b), DtX(br, b);
            X.upload(k->X, d.st);
            c->UtX.alloc(n, b);
            d.gemm(1.0, c->U1, X, 0.0, c->UtX);
            c->DtXt.alloc(b, br);
            d.gemm(1.0, c->DU, X, 0.0, DtX);                 // Design_U'X (fixed-Lambda sweep only)
            d.transpose(DtX, c->DtXt);
            d.sync();
        }
        d.sync();
    }
    // accuracy of the diagonalisation identities the "diag" forms rely on (SURVEY.md section 8d caveat)
    c->M2.alloc(br, br);
    d.gemm(1.0, c->DU, c->DU, 0.0, c->M2);                      // Design_U' Design_U  ~ diag(s2)
    unsigned long long* d_mx = nullptr;
    BSFG_CUDA(cudaMalloc(&d_mx, 2 * sizeof(unsigned long long)));
    BSFG_CUDA(cudaMemsetAsync(d_mx, 0, 2 * sizeof(unsigned long long), d.st));
    diag_residual_kernel<<<d.sms * 4, 256, 0, d.st>>>(c->M2.p, c->M2.ld, br, c->s2_2.p, d_mx + 1);
    BSFG_CHECK_LAUNCH(); g_launch_counter++;
    bool have_m1 = false;
    if (k->Ainv) {
        DMat Ainv(r, r), P(br, br), PU(br, br);
        Ainv.upload(k->Ainv, d.st);
        build_prior_prec_kernel<<<d.sms * 4, 256, 0, d.st>>>(Ainv.p, Ainv.ld, r, b, c->pri.b_X_prec, P.p, P.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        c->M1.alloc(br, br);
        d.gemm(1.0, P, c->U2, 0.0, PU);                         // P symmetric: P'U2 = P U2
        d.gemm(1.0, c->U2, PU, 0.0, c->M1);                     // U2' P U2  ~ diag(s1)
        diag_residual_kernel<<<d.sms * 4, 256, 0, d.st>>>(c->M1.p, c->M1.ld, br, c->s1_2.p, d_mx);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        d.sync();
        have_m1 = true;
    }
    unsigned long long h_mx[2];
    BSFG_CUDA(cudaMemcpyAsync(h_mx, d_mx, sizeof h_mx, cudaMemcpyDeviceToHost, d.st));
    d.sync();
    cudaFree(d_mx);
    double r1, r2;
    memcpy(&r1, &h_mx[0], 8); memcpy(&r2, &h_mx[1], 8);
    c->resid_s1 = have_m1 ? r1 : -1.0;
    c->resid_s2 = r2;
    // the residuals are judged relative to the size of the diagonal they are compared with
    double tol1 = 1e-11, tol2 = 1e-11;
    {
        std::vector<double> h1(br), h2(br);
        c->s1_2.download(h1.data(), d.st); c->s2_2.download(h2.data(), d.st);
        d.sync();
        double m1 = 1.0, m2 = 1.0;
        for (int i = 0; i < br; i++) { m1 = std::max(m1, fabs(h1[i])); m2 = std::max(m2, fabs(h2[i])); }
        tol1 *= m1; tol2 *= m2;
    }
    int mode = c->cfg.mode;
    if (mode == 0) {
        if (have_m1) mode = (r1 <= tol1 && r2 <= tol2) ? 2 : 1;
        else if (r2 <= tol2) mode = 2;
        else fail(BSFG_EARG, "Design_U'Design_U differs from diag(s2) by %.3g (tolerance %.3g): the diagonal shortcuts for "
                             "resid_Y_prec / E_a_prec would be inaccurate; pass Ainv in bsfg_constants so the direct forms can be used", r2, tol2);
    }
    if (mode == 1 && !have_m1) fail(BSFG_EARG, "mode 'direct' needs Ainv");
    c->mode = mode;
    if (mode == 1) { c->M1th.alloc(br, pl); c->M2th.alloc(br, pl); }
    else { c->M1.release(); c->M2.release(); }
    c->have_consts = true;
    BSFG_API_END
}

extern "C" int bsfg_get_mode(bsfg_ctx* c, int* mode, double* resid_s1, double* resid_s2) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c);
    if (mode) *mode = c->mode;
    if (resid_s1) *resid_s1 = c->resid_s1;
    if (resid_s2) *resid_s2 = c->resid_s2;
    BSFG_API_END
}

extern "C" int bsfg_get_gram_stats(bsfg_ctx* c, int* mode, int* terms, double* step, long long* fallbacks,
                                   double* c_min, double* c_max, double* ms_nodes, double* ms_traits) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c);
    if (mode) *mode = c->gram.mode;
    if (terms) *terms = c->gram.terms;
    if (step) *step = c->gram.step;
    ExpSumPlan pl;
    memset(&pl, 0, sizeof pl);
    unsigned long long fb = 0;
    if (c->lsc.plan) {
        BSFG_CUDA(cudaMemcpyAsync(&pl, c->lsc.plan, sizeof pl, cudaMemcpyDeviceToHost, c->dev.st));
        BSFG_CUDA(cudaMemcpyAsync(&fb, c->lsc.fallbacks, sizeof fb, cudaMemcpyDeviceToHost, c->dev.st));
        c->dev.sync();
    }
    if (fallbacks) *fallbacks = (long long)fb;
    if (c_min) *c_min = pl.c_min;
    if (c_max) *c_max = pl.c_max;
    // split of the last timed iteration's "lambda" phase (expsum mode): G'A over the n individuals, then (G'A)'E over the traits
    float t_nodes = 0, t_traits = 0;
    if (c->lsc.sub_timed && c->gram.mode == GRAM_EXPSUM) {
        c->dev.sync();
        BSFG_CUDA(cudaEventElapsedTime(&t_nodes, c->ev_phase[1], c->lsc.ev_sub[0]));
        BSFG_CUDA(cudaEventElapsedTime(&t_traits, c->lsc.ev_sub[0], c->lsc.ev_sub[1]));
    }
    if (ms_nodes) *ms_nodes = t_nodes;
    if (ms_traits) *ms_traits = t_traits;
    BSFG_API_END
}

extern "C" int bsfg_set_state(bsfg_ctx* c, const bsfg_state* st) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c); REQUIRE_PTR(st);
    // replicated members by broadcast from rank 0 (bsfg_state::bcast_replicated): only rank 0 has to supply them
    const bool bcast = st->bcast_replicated != 0 && c->cfg.world > 1;
    const bool give = !bcast || c->cfg.rank == 0;
    if (bcast && !c->comm) fail(BSFG_ESTATE, "bcast_replicated needs bsfg_comm_init");
    if (give) { REQUIRE_PTR(st->delta); REQUIRE_PTR(st->tauh); }
    REQUIRE_PTR(st->resid_Y_prec); REQUIRE_PTR(st->E_a_prec); REQUIRE_PTR(st->Plam);
    if (c->b > 0) REQUIRE_PTR(st->B);
    if (c->has_missing) { REQUIRE_PTR(st->Lambda); REQUIRE_PTR(st->E_a); }     // the imputation step conditions on them (R:181)
    const int k = st->k;
    if (k < 2 || k > c->kmax) fail(BSFG_ESHAPE, "bsfg_set_state: k = %d outside [2, k_max = %d]", k, c->kmax);
    Device& d = c->dev;
    const int n = c->n, pl = c->pl, r = c->r, b = c->b;
    c->k = k;
    c->iter = st->nrun;
    int px_flag = (st->px_factor != nullptr && st->F_px != nullptr) ? 1 : 0;
    if (bcast) {    // every rank learns from rank 0 whether this is a px chain and whether F_a / F_h2 come along
        int flags[3] = {px_flag, st->F_a ? 1 : 0, st->F_h2 ? 1 : 0};
        int* dflags = reinterpret_cast<int*>(c->colsum_cnt.p);       // 2 k_max doubles of scratch: room for three ints
        if (c->cfg.rank == 0) BSFG_CUDA(cudaMemcpyAsync(dflags, flags, sizeof flags, cudaMemcpyHostToDevice, d.st));
        BSFG_NCCL(g_nccl.Broadcast(dflags, dflags, sizeof flags, ncclChar, 0, c->comm, d.st));
        BSFG_CUDA(cudaMemcpyAsync(flags, dflags, sizeof flags, cudaMemcpyDeviceToHost, d.st));
        d.sync();
        px_flag = flags[0];
        c->bc_have_fa = flags[1] != 0; c->bc_have_h2 = flags[2] != 0;
    }
    c->px_mode = px_flag != 0;
    DMat F = view(c->F, n, k);
    if (give) {
        if (c->px_mode) {      // parameter-expanded state: the device keeps F_px (F = F_px / sqrt(px_factor) is derived on get_state)
            F.upload(st->F_px, d.st);
            BSFG_CUDA(cudaMemcpyAsync(c->px.p, st->px_factor, sizeof(double) * k, cudaMemcpyHostToDevice, d.st));
        } else {
            REQUIRE_PTR(st->F);
            F.upload(st->F, d.st);
        }
    }
    c->rp.upload(st->resid_Y_prec, d.st); c->ep.upload(st->E_a_prec, d.st);
    if (give) {
        BSFG_CUDA(cudaMemcpyAsync(c->delta.p, st->delta, sizeof(double) * k, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(c->tauh.p, st->tauh, sizeof(double) * k, cudaMemcpyHostToDevice, d.st));
    }
    if (b > 0) c->Bfix.upload(st->B, d.st);
    {
        DMat Plam = view(c->Ptmp, pl, k), Plamt = view(c->Plamt, k, pl);   // persistent scratch: no cudaMalloc on the state path
        Plam.upload(st->Plam, d.st);
        d.transpose(Plam, Plamt);
    }
    if (st->Lambda) {
        DMat Lam = view(c->Lam, pl, k), Lt = view(c->Lt, k, pl);
        Lam.upload(st->Lambda, d.st);
        if (c->px_mode) {   // the device keeps Lambda_px = Lambda / sqrt(px_factor)
            px_vectors_kernel<<<ceil_div(k, 128), 128, 0, d.st>>>(c->px.p, k, c->px_sq.p, c->px_isq.p, c->px_ipx.p);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
            d.scale(Lam, nullptr, c->px_isq.p, Lam);
        }
        d.transpose(Lam, Lt);
    }
    if (st->Lambda_prec) { DMat LP = view(c->LP, pl, k); LP.upload(st->Lambda_prec, d.st); }
    if (give && st->F_a) { DMat Fa = view(c->Fa, r, k); Fa.upload(st->F_a, d.st); }
    if (give && st->F_h2) BSFG_CUDA(cudaMemcpyAsync(c->F_h2.p, st->F_h2, sizeof(double) * k, cudaMemcpyHostToDevice, d.st));
    if (bcast) {
        auto bc = [&](double* p, size_t count) { BSFG_NCCL(g_nccl.Broadcast(p, p, count, ncclDouble, 0, c->comm, d.st)); };
        bc(c->F.p, (size_t)c->F.ld * (size_t)k);
        if (c->bc_have_fa) bc(c->Fa.p, (size_t)c->Fa.ld * (size_t)k);
        if (c->bc_have_h2) bc(c->F_h2.p, (size_t)k);
        bc(c->delta.p, (size_t)k); bc(c->tauh.p, (size_t)k);
        if (c->px_mode) bc(c->px.p, (size_t)k);
    }
    if (st->E_a) {
        if (c->Ea_in.rows != r) c->Ea_in.alloc(r, pl);
        c->Ea_in.upload(st->E_a, d.st);
    }
    if (c->r2 > 0) {
        REQUIRE_PTR(st->W); REQUIRE_PTR(st->W_prec);
        c->W.upload(st->W, d.st); c->wp.upload(st->W_prec, d.st);
        d.transpose(c->W, c->Wt);
    }
    c->theta_valid = false;
    c->derived_valid = false;
    d.sync();
    c->have_state = true;
    BSFG_API_END
}

extern "C" int bsfg_get_state(bsfg_ctx* c, bsfg_state* st) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c); REQUIRE_PTR(st);
    if (!c->have_state) fail(BSFG_ESTATE, "no state on the device");
    Device& d = c->dev;
    const int n = c->n, pl = c->pl, r = c->r, b = c->b, br = c->br, k = c->k;
    st->k = k;
    st->nrun = (int)c->iter;
    static const bool trace = getenv("BSFG_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        d.sync();
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[bsfg_get_state] %-12s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    if (c->px_mode) {
        px_vectors_kernel<<<ceil_div(k, 128), 128, 0, d.st>>>(c->px.p, k, c->px_sq.p, c->px_isq.p, c->px_ipx.p);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    }
    if (st->Lambda) {
        DMat Lam = view(c->Lam, pl, k);
        if (c->px_mode) {   // Lambda = Lambda_px * sqrt(px_factor), fast_BSFG_sampler_px.R:267
            DMat tmp = view(c->Ptmp, pl, k);
            d.scale(Lam, nullptr, c->px_sq.p, tmp);
            tmp.download(st->Lambda, d.st);
            d.sync();
        } else Lam.download(st->Lambda, d.st);
    }
    lap("Lambda");
    if (st->Lambda_prec) { DMat LP = view(c->LP, pl, k); LP.download(st->Lambda_prec, d.st); }
    lap("Lambda_prec");
    if (st->Plam) {
        DMat Plam = view(c->Ptmp, pl, k), Plamt = view(c->Plamt, k, pl);
        d.transpose(Plamt, Plam);
        Plam.download(st->Plam, d.st);
    }
    lap("Plam");
    if (st->F) {
        DMat F = view(c->F, n, k);
        if (c->px_mode) {   // F = F_px / sqrt(px_factor), fast_BSFG_sampler_px.R:294
            DMat tmp = view(c->YL, n, k);
            d.scale(F, nullptr, c->px_isq.p, tmp);
            tmp.download(st->F, d.st);
            d.sync();
        } else F.download(st->F, d.st);
    }
    if (c->px_mode) {
        if (st->F_px) { DMat F = view(c->F, n, k); F.download(st->F_px, d.st); }
        if (st->px_factor) BSFG_CUDA(cudaMemcpyAsync(st->px_factor, c->px.p, sizeof(double) * k, cudaMemcpyDeviceToHost, d.st));
    }
    if (st->F_a) { DMat Fa = view(c->Fa, r, k); Fa.download(st->F_a, d.st); }
    if (st->F_h2) BSFG_CUDA(cudaMemcpyAsync(st->F_h2, c->F_h2.p, sizeof(double) * k, cudaMemcpyDeviceToHost, d.st));
    if (st->B && b > 0) c->Bfix.download(st->B, d.st);
    lap("F,F_a,h2,B");
    if (st->E_a) {
        if (c->theta_valid) {
            // E_a = location_sample[b+(1:r),] = U2[b.., :] theta   (fast_BSFG_sampler.R:207); U2t columns b.. = that block'
            DMat Ea = view(c->tmp_brp, r, pl);       // scratch that is dead between iterations
            DMat U2Et = view_at(c->U2t.p + (long)b * c->U2t.ld, br, r, c->U2t.ld);
            d.gemm(1.0, U2Et, c->theta, 0.0, Ea);
            Ea.download(st->E_a, d.st);
        } else if (c->Ea_in.p) {
            c->Ea_in.download(st->E_a, d.st);
        } else {
            fail(BSFG_ESTATE, "E_a requested but neither a sweep nor set_state provided it");
        }
    }
    if (c->r2 > 0) {
        if (st->W) c->W.download(st->W, d.st);
        if (st->W_prec) c->wp.download(st->W_prec, d.st);
    }
    if (st->resid_Y_prec) c->rp.download(st->resid_Y_prec, d.st);
    if (st->E_a_prec) c->ep.download(st->E_a_prec, d.st);
    if (st->delta) BSFG_CUDA(cudaMemcpyAsync(st->delta, c->delta.p, sizeof(double) * k, cudaMemcpyDeviceToHost, d.st));
    if (st->tauh) BSFG_CUDA(cudaMemcpyAsync(st->tauh, c->tauh.p, sizeof(double) * k, cudaMemcpyDeviceToHost, d.st));
    d.sync();
    lap("rest");
    BSFG_API_END
}

extern "C" int bsfg_get_aux(bsfg_ctx* c, bsfg_aux* a) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c); REQUIRE_PTR(a);
    Device& d = c->dev;
    const int pl = c->pl, k = c->k;
    if (a->Lambda_prec_rate) BSFG_CUDA(cudaMemcpyAsync(a->Lambda_prec_rate, c->aux_lp_rate.p, sizeof(double) * (size_t)pl * k, cudaMemcpyDeviceToHost, d.st));
    if (a->resid_Y_prec_rate) c->aux_rr.download(a->resid_Y_prec_rate, d.st);
    if (a->E_a_prec_rate) c->aux_re.download(a->E_a_prec_rate, d.st);
    if (a->delta_rate) BSFG_CUDA(cudaMemcpyAsync(a->delta_rate, c->delta_rates.p, sizeof(double) * (k + 2), cudaMemcpyDeviceToHost, d.st));
    if (a->W_prec_rate && c->r2 > 0) c->aux_rw.download(a->W_prec_rate, d.st);
    if (a->Y_imputed && c->has_missing) c->Ycur.download(a->Y_imputed, d.st);
    if (a->px_rate && c->px_mode) BSFG_CUDA(cudaMemcpyAsync(a->px_rate, c->aux_px_rate.p, sizeof(double) * k, cudaMemcpyDeviceToHost, d.st));
    if (a->h2_log_ps) BSFG_CUDA(cudaMemcpyAsync(a->h2_log_ps, c->aux_logps.p, sizeof(double) * (size_t)k * c->H, cudaMemcpyDeviceToHost, d.st));
    d.sync();
    BSFG_API_END
}

static int sweep_impl(bsfg_ctx* c, int n_iter, bool px) {
    BSFG_API_BEGIN
    require_ready(c);
    if (n_iter < 0) fail(BSFG_EARG, "n_iter < 0");
    if (px != c->px_mode)
        fail(BSFG_ESTATE, px ? "bsfg_sweep_px: the state was set without px_factor / F_px"
                             : "bsfg_sweep: the state is parameter-expanded (px_factor / F_px were set); use bsfg_sweep_px");
    Device& d = c->dev;
    c->aux_on = false;
    if (c->gram.auto_terms && c->lsc.plan) {
        // re-choose the number of exponential-sum nodes from the last plan (identical on every rank: the c range is
        // max-allreduced before the plan is made), with a 4x margin on both ends of the range
        ExpSumPlan pl;
        BSFG_CUDA(cudaMemcpyAsync(&pl, c->lsc.plan, sizeof pl, cudaMemcpyDeviceToHost, d.st));
        d.sync();
        if (pl.R > 0 && pl.c_min > 0.0 && std::isfinite(pl.c_max)) {
            int need = expsum_terms_needed(c->gram.s_min, c->gram.s_max, pl.c_min / 4.0, pl.c_max * 4.0, c->gram.step);
            need = (int)round_up(std::max(need, 64), 16);
            c->gram.terms = std::min(need, c->gram.terms_max);
        }
    }
    const long long launches_before = g_launch_counter;
    BSFG_CUDA(cudaEventRecord(c->ev_begin, d.st));
    for (int it = 0; it < n_iter; it++) {
        const long i = c->iter + 1;                                         // for(i in start_i+(1:n_samples)), R:176
        const uint64_t seed = c->cfg.seed;
        const uint32_t ii = (uint32_t)i;
        IterVariates v;
        v.zL = make_var(nullptr, seed, SITE_Z_LAMBDA, ii, c->p_off, c->k);
        v.zM = make_var(nullptr, seed, SITE_Z_MEANS, ii, c->p_off, round_up(c->br, 2));    // even: a row pair = one Philox pair
        v.uh2 = make_var(nullptr, seed, SITE_U_H2, ii, 0, c->k);
        v.zFa = make_var(nullptr, seed, SITE_Z_FA, ii, 0, c->r);
        v.zF = make_var(nullptr, seed, SITE_Z_F, ii, 0, c->n);
        v.gLP = make_var(nullptr, seed, SITE_G_LAMBDA_PREC, ii, c->p_off, c->p_total);
        v.gR = make_var(nullptr, seed, SITE_G_RESID, ii, c->p_off, 1);
        v.gE = make_var(nullptr, seed, SITE_G_EA, ii, c->p_off, 1);
        v.gD = make_var(nullptr, seed, SITE_G_DELTA, ii, 0, 1);
        v.zW = make_var(nullptr, seed, SITE_Z_MEANS_W, ii, c->p_off, c->r2 > 0 ? round_up(c->r2, 2) : 2);
        v.gW = make_var(nullptr, seed, SITE_G_W, ii, c->p_off, 1);
        v.zY = make_var(nullptr, seed, SITE_Z_Y, ii, c->p_off, c->n);
        v.gPX = make_var(nullptr, seed, SITE_G_PX, ii, 0, 1);
        c->phase_timing = (it == n_iter - 1);
        // The phase-timed iteration of a multi-iteration call runs on ONE stream: with the side stream on, its kernels land inside
        // whatever phase happens to be running (the F_a products inside the trait product, say) and the CUDA-event intervals
        // no longer measure the kernels they bracket.  One iteration in n_iter pays ~3% for clean phase times; single-iteration
        // calls (the end-to-end path) keep the overlap.
        c->serial_iter = c->phase_timing && n_iter > 1;
        // BSFG_PROFILE_ITER=<n>: bracket the n-th Gibbs iteration of this process (0-based) with cudaProfilerStart/Stop, so that
        // `ncu --profile-from-start off` captures exactly one warm iteration (tools/profile_iteration.sh)
        static const long prof_iter = [] { const char* e = getenv("BSFG_PROFILE_ITER"); return e ? atol(e) : -1L; }();
        static long iters_run = 0;
        const bool prof = prof_iter >= 0 && iters_run == prof_iter;
        if (prof) { d.sync(); cudaProfilerStart(); }
        gibbs_iteration(c, v, false, px);
        if (prof) { d.sync(); if (c->side_ready) cudaStreamSynchronize(c->dev2.st); cudaProfilerStop(); }
        iters_run++;
        RngStream rs{seed, SITE_U_K, ii};
        c->iter = i;             // the Gibbs body of iteration i has run: an update_k failure must not leave the chain half-stepped
        update_k_step(c, i, philox_uniform(rs, 0), nullptr);               // uu = runif(1) every iteration, R:325
    }
    finish_timing(c, launches_before, n_iter);
    BSFG_API_END
}
extern "C" int bsfg_sweep(bsfg_ctx* c, int n_iter) { return sweep_impl(c, n_iter, false); }
extern "C" int bsfg_sweep_px(bsfg_ctx* c, int n_iter) { return sweep_impl(c, n_iter, true); }

namespace bsfg {
// one saved sample: save_posterior_samples(sp_num, ...)   BSFG_functions.R:375-408
static void posterior_store(bsfg_ctx* c, int sp_num) {
    if (c->post_n >= c->post_cap) fail(BSFG_ESTATE, "posterior capacity (%d samples) exhausted: call bsfg_posterior_begin with more", c->post_cap);
    if (!c->theta_valid) fail(BSFG_ESTATE, "no sweep has run since set_state: nothing to store");
    Device& d = c->dev;
    const int n = c->n, pl = c->pl, r = c->r, k = c->k, km = c->kmax, s = c->post_n;
    auto store = [&](const double* cur, long ld, long rows, DMat& tr) {
        const long total = rows * km;
        trace_store_kernel<<<(int)std::min<long>((total + 255) / 256, (long)d.sms * 8), 256, 0, d.st>>>(cur, ld, rows, k, km,
                                                                                                 tr.p + (long)s * tr.ld);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
    };
    if (c->px_mode) {
        // the px loop saves Lambda = Lambda_px sqrt(px) and F = F_px / sqrt(px)   (fast_BSFG_sampler_px.R:267, 294, 317)
        px_vectors_kernel<<<ceil_div(k, 128), 128, 0, d.st>>>(c->px.p, k, c->px_sq.p, c->px_isq.p, c->px_ipx.p);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        DMat Lam = view(c->Lam, pl, k), Ls = view(c->Ptmp, pl, k), F = view(c->F, n, k), Fs = view(c->YL, n, k);
        d.scale(Lam, nullptr, c->px_sq.p, Ls);
        d.scale(F, nullptr, c->px_isq.p, Fs);
        store(Ls.p, Ls.ld, pl, c->tr_Lambda);
        store(Fs.p, Fs.ld, n, c->tr_F);
    } else {
        store(c->Lam.p, c->Lam.ld, pl, c->tr_Lambda);
        store(c->F.p, c->F.ld, n, c->tr_F);
    }
    store(c->Fa.p, c->Fa.ld, r, c->tr_Fa);
    store(c->delta.p, 1, 1, c->tr_delta);
    store(c->F_h2.p, 1, 1, c->tr_h2);
    BSFG_CUDA(cudaMemcpyAsync(c->tr_rp.p + (long)s * c->tr_rp.ld, c->rp.p, sizeof(double) * pl, cudaMemcpyDeviceToDevice, d.st));
    BSFG_CUDA(cudaMemcpyAsync(c->tr_ep.p + (long)s * c->tr_ep.ld, c->ep.p, sizeof(double) * pl, cudaMemcpyDeviceToDevice, d.st));
    if (c->r2 > 0)
        BSFG_CUDA(cudaMemcpyAsync(c->tr_wp.p + (long)s * c->tr_wp.ld, c->wp.p, sizeof(double) * pl, cudaMemcpyDeviceToDevice, d.st));
    // B and E_a are linear in theta ([B;E_a] = U theta, cpp:78), so their running means are U * (running mean of theta)
    {
        const long total = (long)c->br * pl;
        mean_update_kernel<<<(int)std::min<long>((total + 255) / 256, (long)d.sms * 8), 256, 0, d.st>>>(
            c->mean_theta.p, c->mean_theta.ld, c->theta.p, c->theta.ld, c->br, pl, (double)sp_num);
        BSFG_CHECK_LAUNCH(); g_launch_counter++;
        if (c->r2 > 0) {
            const long tw = (long)c->r2 * pl;
            mean_update_kernel<<<(int)std::min<long>((tw + 255) / 256, (long)d.sms * 8), 256, 0, d.st>>>(
                c->mean_W.p, c->mean_W.ld, c->W.p, c->W.ld, c->r2, pl, (double)sp_num);
            BSFG_CHECK_LAUNCH(); g_launch_counter++;
        }
    }
    c->post_spnum.push_back(sp_num); c->post_k.push_back(k);
    c->post_n++;
}
}  // namespace bsfg

extern "C" int bsfg_posterior_begin(bsfg_ctx* c, int capacity) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c);
    if (capacity <= 0) fail(BSFG_EARG, "capacity must be positive");
    const int n = c->n, pl = c->pl, r = c->r, km = c->kmax;
    c->tr_Lambda.alloc((long)pl * km, capacity); c->tr_F.alloc((long)n * km, capacity); c->tr_Fa.alloc((long)r * km, capacity);
    c->tr_delta.alloc(km, capacity); c->tr_h2.alloc(km, capacity);
    c->tr_rp.alloc(pl, capacity); c->tr_ep.alloc(pl, capacity);
    if (c->r2 > 0) { c->tr_wp.alloc(pl, capacity); c->mean_W.alloc(c->r2, pl); }
    c->mean_theta.alloc(c->br, pl);
    c->post_cap = capacity; c->post_n = 0; c->post_spnum.clear(); c->post_k.clear();
    BSFG_API_END
}

extern "C" int bsfg_posterior_count(bsfg_ctx* c, int* n_samples, int* k_max) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c);
    if (n_samples) *n_samples = c->post_n;
    if (k_max) *k_max = c->kmax;
    BSFG_API_END
}

extern "C" int bsfg_get_posterior(bsfg_ctx* c, bsfg_posterior* o) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c); REQUIRE_PTR(o);
    Device& d = c->dev;
    const int S = c->post_n, pl = c->pl, r = c->r, b = c->b, br = c->br;
    auto dl = [&](const DMat& tr, double* host) {
        if (!host || S == 0) return;
        BSFG_CUDA(cudaMemcpy2DAsync(host, tr.rows * sizeof(double), tr.p, tr.ld * sizeof(double), tr.rows * sizeof(double), S,
                                    cudaMemcpyDeviceToHost, d.st));
    };
    dl(c->tr_Lambda, o->Lambda); dl(c->tr_F, o->F); dl(c->tr_Fa, o->F_a); dl(c->tr_delta, o->delta); dl(c->tr_h2, o->F_h2);
    dl(c->tr_rp, o->resid_Y_prec); dl(c->tr_ep, o->E_a_prec);
    if (c->r2 > 0) { dl(c->tr_wp, o->W_prec); if (o->W && S > 0) c->mean_W.download(o->W, d.st); }
    if (S > 0 && (o->B || o->E_a)) {
        DMat loc = view(c->tmp_brp, br, pl);                    // scratch that is dead between iterations
        d.gemm(1.0, c->U2t, c->mean_theta, 0.0, loc);           // mean([B;E_a]) = U mean(theta)
        if (o->B && b > 0)
            BSFG_CUDA(cudaMemcpy2DAsync(o->B, b * sizeof(double), loc.p, loc.ld * sizeof(double), b * sizeof(double), pl,
                                        cudaMemcpyDeviceToHost, d.st));
        if (o->E_a)
            BSFG_CUDA(cudaMemcpy2DAsync(o->E_a, r * sizeof(double), loc.p + b, loc.ld * sizeof(double), r * sizeof(double), pl,
                                        cudaMemcpyDeviceToHost, d.st));
    }
    if (o->sp_num) for (int s = 0; s < S; s++) o->sp_num[s] = c->post_spnum[s];
    if (o->k) for (int s = 0; s < S; s++) o->k[s] = c->post_k[s];
    d.sync();
    BSFG_API_END
}

extern "C" int bsfg_sweep_collect(bsfg_ctx* c, int n_iter, int burn, int thin) {
    BSFG_API_BEGIN
    require_ready(c);
    if (n_iter < 0 || thin <= 0 || burn < 0) fail(BSFG_EARG, "bsfg_sweep_collect: n_iter >= 0, burn >= 0, thin > 0 required");
    if (c->post_cap <= 0) fail(BSFG_ESTATE, "call bsfg_posterior_begin first");
    int done = 0;
    while (done < n_iter) {
        // run up to the next thinning boundary (fast_BSFG_sampler.R:272), store, continue
        long i = c->iter, nxt = std::max<long>(i + 1, burn + 1);
        while ((nxt - burn) % thin != 0) nxt++;
        const int step = (int)std::min<long>(nxt - i, n_iter - done);
        const int rc = c->px_mode ? bsfg_sweep_px(c, step) : bsfg_sweep(c, step);
        if (rc != BSFG_OK) return rc;
        done += step;
        if (c->iter == nxt) posterior_store(c, (int)((nxt - burn) / thin));
    }
    c->dev.sync();
    BSFG_API_END
}

static int sweep_injected_impl(bsfg_ctx* c, const bsfg_variates* hv, bool px) {
    BSFG_API_BEGIN
    require_ready(c);
    REQUIRE_PTR(hv);
    if (px != c->px_mode) fail(BSFG_ESTATE, "sweep_injected: px mode of the call and of the state differ");
    Device& d = c->dev;
    const int n = c->n, pl = c->pl, k = c->k, r = c->r, br = c->br;
    const long i = c->iter + 1;
    const uint64_t seed = c->cfg.seed;
    const uint32_t ii = (uint32_t)i;
    auto stage = [&](DMat& buf, const double* host, long rows, long cols) -> const DMat* {
        if (!host) return nullptr;
        if (buf.rows != rows || buf.cols != cols) buf.alloc(rows, cols);
        buf.upload(host, d.st);
        return &buf;
    };
    IterVariates v;
    v.zL = make_var(stage(c->v_zL, hv->z_Lambda, k, pl), seed, SITE_Z_LAMBDA, ii, c->p_off, k);
    v.zM = make_var(stage(c->v_zM, hv->z_means, br, pl), seed, SITE_Z_MEANS, ii, c->p_off, br);
    v.uh2 = make_var(stage(c->v_uh2, hv->u_h2, k, 1), seed, SITE_U_H2, ii, 0, k);
    v.zFa = make_var(stage(c->v_zFa, hv->z_Fa, r, k), seed, SITE_Z_FA, ii, 0, r);
    v.zF = make_var(stage(c->v_zF, hv->z_F, n, k), seed, SITE_Z_F, ii, 0, n);
    v.gLP = make_var(stage(c->v_gLP, hv->g_Lambda_prec, pl, k), seed, SITE_G_LAMBDA_PREC, ii, c->p_off, c->p_total);
    v.gR = make_var(stage(c->v_gR, hv->g_resid, pl, 1), seed, SITE_G_RESID, ii, c->p_off, 1);
    v.gE = make_var(stage(c->v_gE, hv->g_Ea, pl, 1), seed, SITE_G_EA, ii, c->p_off, 1);
    v.zW = make_var(c->r2 > 0 ? stage(c->v_zW, hv->z_W, c->r2, pl) : nullptr, seed, SITE_Z_MEANS_W, ii, c->p_off, c->r2 > 0 ? c->r2 : 1);
    v.gW = make_var(c->r2 > 0 ? stage(c->v_gW, hv->g_W, pl, 1) : nullptr, seed, SITE_G_W, ii, c->p_off, 1);
    v.zY = make_var(c->has_missing ? stage(c->v_zY, hv->z_Y, n, pl) : nullptr, seed, SITE_Z_Y, ii, c->p_off, n);
    v.gPX = make_var(px ? stage(c->v_gpx, hv->g_px, k, 1) : nullptr, seed, SITE_G_PX, ii, 0, 1);
    {
        const DMat* gd = nullptr;
        if (hv->g_delta) {       // only as many entries as sample_delta draws: 1 + |2:(k-1)|
            if (c->v_gD.rows != c->kmax + 2) c->v_gD.alloc(c->kmax + 2, 1);
            BSFG_CUDA(cudaMemcpyAsync(c->v_gD.p, hv->g_delta, sizeof(double) * (size_t)(k == 2 ? 3 : k - 1),
                                      cudaMemcpyHostToDevice, d.st));
            gd = &c->v_gD;
        }
        v.gD = make_var(gd, seed, SITE_G_DELTA, ii, 0, 1);
    }
    d.sync();
    c->aux_on = true;
    c->phase_timing = true;
    c->serial_iter = false;
    const long long launches_before = g_launch_counter;
    BSFG_CUDA(cudaEventRecord(c->ev_begin, d.st));
    gibbs_iteration(c, v, false, px);
    double uu;
    if (hv->u_k) uu = *hv->u_k;
    else { RngStream rs{seed, SITE_U_K, ii}; uu = philox_uniform(rs, 0); }
    KAddVariates ka;
    bool have_add = hv->k_g_lp && hv->k_g_delta && hv->k_z_lambda && hv->k_u_h2 && hv->k_z_Fa && hv->k_z_F && (!px || hv->k_g_px);
    if (have_add) {
        // staged as one buffer: [g_lp (pl) | z_lambda (pl) | z_Fa (r) | z_F (n)]
        const long tot = 2L * pl + r + n;
        if (c->v_kadd.rows != tot) c->v_kadd.alloc(tot, 1);
        BSFG_CUDA(cudaMemcpyAsync(c->v_kadd.p, hv->k_g_lp, sizeof(double) * pl, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(c->v_kadd.p + pl, hv->k_z_lambda, sizeof(double) * pl, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(c->v_kadd.p + 2L * pl, hv->k_z_Fa, sizeof(double) * r, cudaMemcpyHostToDevice, d.st));
        BSFG_CUDA(cudaMemcpyAsync(c->v_kadd.p + 2L * pl + r, hv->k_z_F, sizeof(double) * n, cudaMemcpyHostToDevice, d.st));
        d.sync();
        auto mk = [&](double* p) { VarSrc s = make_var(nullptr, 0, 0, 0, 0, 1); s.ptr = p; s.ld = 0; return s; };
        ka.g_lp = mk(c->v_kadd.p); ka.z_lam = mk(c->v_kadd.p + pl); ka.z_fa = mk(c->v_kadd.p + 2L * pl);
        ka.z_f = mk(c->v_kadd.p + 2L * pl + r);
        ka.g_delta = *hv->k_g_delta; ka.u_h2 = *hv->k_u_h2; ka.g_px = px ? *hv->k_g_px : 1.0; ka.have_scalars = true;
    }
    c->iter = i;
    update_k_step(c, i, uu, have_add ? &ka : nullptr);
    finish_timing(c, launches_before, 1);
    BSFG_API_END
}
extern "C" int bsfg_sweep_injected(bsfg_ctx* c, const bsfg_variates* hv) { return sweep_injected_impl(c, hv, false); }
extern "C" int bsfg_sweep_px_injected(bsfg_ctx* c, const bsfg_variates* hv) { return sweep_injected_impl(c, hv, true); }

extern "C" int bsfg_sweep_fixed_lambda(bsfg_ctx* c, const double* Lambda, int k, const bsfg_variates* hv) {
    BSFG_API_BEGIN
    require_ready(c);
    REQUIRE_PTR(Lambda);
    if (k != c->k) fail(BSFG_ESHAPE, "bsfg_sweep_fixed_lambda: Lambda has %d columns, the state has k = %d", k, c->k);
    if (c->px_mode) fail(BSFG_ESTATE, "bsfg_sweep_fixed_lambda: the state is parameter-expanded; set a plain state first");
    Device& d = c->dev;
    const int n = c->n, pl = c->pl, r = c->r, br = c->br;
    const long i = c->iter + 1;
    const uint64_t seed = c->cfg.seed;
    const uint32_t ii = (uint32_t)i;
    {   // Lambda = matrix(BSFG_state$Lambda[,j], nr = p)      fast_BSFG_sampler_fixedlambda.R:163-164
        DMat Lam = view(c->Lam, pl, k), Lt = view(c->Lt, k, pl);
        Lam.upload(Lambda, d.st);
        d.transpose(Lam, Lt);
    }
    static const bsfg_variates none = {};
    if (!hv) hv = &none;
    auto stage = [&](DMat& buf, const double* host, long rows, long cols) -> const DMat* {
        if (!host) return nullptr;
        if (buf.rows != rows || buf.cols != cols) buf.alloc(rows, cols);
        buf.upload(host, d.st);
        return &buf;
    };
    IterVariates v;
    v.zL = make_var(nullptr, seed, SITE_Z_LAMBDA, ii, c->p_off, k);
    v.gLP = make_var(nullptr, seed, SITE_G_LAMBDA_PREC, ii, c->p_off, c->p_total);
    v.gD = make_var(nullptr, seed, SITE_G_DELTA, ii, 0, 1);
    v.zM = make_var(stage(c->v_zM, hv->z_means, br, pl), seed, SITE_Z_MEANS, ii, c->p_off, br);
    v.uh2 = make_var(stage(c->v_uh2, hv->u_h2, k, 1), seed, SITE_U_H2, ii, 0, k);
    v.zFa = make_var(stage(c->v_zFa, hv->z_Fa, r, k), seed, SITE_Z_FA, ii, 0, r);
    v.zF = make_var(stage(c->v_zF, hv->z_F, n, k), seed, SITE_Z_F, ii, 0, n);
    v.gR = make_var(stage(c->v_gR, hv->g_resid, pl, 1), seed, SITE_G_RESID, ii, c->p_off, 1);
    v.gE = make_var(stage(c->v_gE, hv->g_Ea, pl, 1), seed, SITE_G_EA, ii, c->p_off, 1);
    v.zW = make_var(c->r2 > 0 ? stage(c->v_zW, hv->z_W, c->r2, pl) : nullptr, seed, SITE_Z_MEANS_W, ii, c->p_off, c->r2 > 0 ? c->r2 : 1);
    v.gW = make_var(c->r2 > 0 ? stage(c->v_gW, hv->g_W, pl, 1) : nullptr, seed, SITE_G_W, ii, c->p_off, 1);
    v.zY = make_var(c->has_missing ? stage(c->v_zY, hv->z_Y, n, pl) : nullptr, seed, SITE_Z_Y, ii, c->p_off, n);
    d.sync();
    c->aux_on = true;
    c->phase_timing = true;
    c->serial_iter = false;
    const long long launches_before = g_launch_counter;
    BSFG_CUDA(cudaEventRecord(c->ev_begin, d.st));
    gibbs_iteration(c, v, /*fixed_lambda=*/true);
    c->iter = i;
    finish_timing(c, launches_before, 1);
    BSFG_API_END
}

extern "C" int bsfg_get_timing(bsfg_ctx* c, double* total_ms, long long* launches, double* phase_ms) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c);
    if (total_ms) *total_ms = c->last_total_ms;
    if (launches) *launches = c->last_launches;
    if (phase_ms) memcpy(phase_ms, c->phase_ms, sizeof(double) * BSFG_N_PHASES);
    BSFG_API_END
}

extern "C" int bsfg_nccl_unique_id(char id_out[128]) {
    BSFG_API_BEGIN
    REQUIRE_PTR(id_out);
    g_nccl.load();
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    BSFG_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, 128);
    BSFG_API_END
}

extern "C" int bsfg_comm_init(bsfg_ctx* c, const char id_in[128]) {
    BSFG_API_BEGIN
    REQUIRE_PTR(c); REQUIRE_PTR(id_in);
    if (c->cfg.world <= 1) return BSFG_OK;
    g_nccl.load();
    ncclUniqueId id;
    memcpy(&id, id_in, 128);
    BSFG_NCCL(g_nccl.CommInitRank(&c->comm, c->cfg.world, id, c->cfg.rank));
    {   // second communicator for the side stream: rank 0 draws its id and hands it out over the first one
        ncclUniqueId id2;
        memset(&id2, 0, sizeof id2);
        if (c->cfg.rank == 0) BSFG_NCCL(g_nccl.GetUniqueId(&id2));
        char* dbuf = nullptr;
        BSFG_CUDA(cudaMalloc(&dbuf, sizeof id2));
        BSFG_CUDA(cudaMemcpyAsync(dbuf, &id2, sizeof id2, cudaMemcpyHostToDevice, c->dev.st));
        BSFG_NCCL(g_nccl.Broadcast(dbuf, dbuf, sizeof id2, ncclChar, 0, c->comm, c->dev.st));
        BSFG_CUDA(cudaMemcpyAsync(&id2, dbuf, sizeof id2, cudaMemcpyDeviceToHost, c->dev.st));
        c->dev.sync();
        cudaFree(dbuf);
        BSFG_NCCL(g_nccl.CommInitRank(&c->comm2, c->cfg.world, id2, c->cfg.rank));
    }
    if (shard_replicated_enabled(2)) {
        c->lsc.rank = c->cfg.rank; c->lsc.world = c->cfg.world;
        c->lsc.allgather = [c](double* buf, size_t chunk) { allgather_inplace(c, buf, chunk); };
        c->lsc.allreduce_max = [c](double* buf, size_t count) {
            BSFG_NCCL(g_nccl.AllReduce(buf, buf, count, ncclDouble, ncclMax, c->comm, c->dev.st));
        };
    }
    BSFG_API_END
}
