// glg_capi.cu -- C-ABI (include/glgym.h) over the sm_100a kernels.  Host side: handle, device buffers, launches.
// No CPU fallback anywhere: every entry point either launches CUDA work or fails with GLG_ERR_CUDA.
#include "../../include/glgym.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <thread>
#include <vector>

#undef GLG_NX
#undef GLG_NU
#undef GLG_ND
#undef GLG_NP
#undef GLG_NINFO
#undef GLG_NSTATS
#include "glg_kernels.cuh"
#include "glg_roles.cuh"
#include "glg_rollout.cuh"


static thread_local std::string g_create_error = "";

struct glg_handle {
    glg_config cfg;
    int B = 0, obs_dim = 0, nt = 64, role_lanes = 32, sms = 148;
    bool have_params = false, have_weather = false, general = false, is_reset = false;
    GlgUniform uni;
    // device buffers
    double *x = nullptr, *u = nullptr, *time = nullptr, *ep_return = nullptr, *ep_info = nullptr;
    int *timestep = nullptr, *table = nullptr, *ep_len = nullptr;
    unsigned int *step_ctr = nullptr;
    float *obs = nullptr, *term_obs = nullptr, *actions = nullptr, *obs_head = nullptr;
    double *reward = nullptr, *info = nullptr, *stats = nullptr;
    unsigned char *done = nullptr, *res_block = nullptr;
    double *weather = nullptr, *start_day = nullptr;
    int *reset_tables = nullptr;
    int n_tables = 0, rows = 0, n_reset_tables = 0;
    // pinned host staging for glg_step_host
    float *h_actions = nullptr, *h_obs = nullptr;
    double *h_reward = nullptr;
    unsigned char *h_done = nullptr;
    // host side of glg_step_host's overlapped observation path (see there)
    std::vector<float> fc_bank;  // [n_tables][rows][5] float32: the weather columns a forecast block shows, as the kernels cast them
    float *h_head = nullptr;     // pinned [B][obs_dim - 5 Np]: staging of the packed columns for a pageable destination
    unsigned char *h_res = nullptr;  // pinned [2 buffers][17 B]: reward | timestep | table | done of every env after a glg_step_host
    int kt_cur = 0;              // buffer the last glg_step_host filled
    bool kt_valid = false;       // h_res[kt_cur] holds this handle's last host step (a prediction source, verified after every step)
    std::vector<int> fc_pred;    // [2][B]: (table, first weather row) the forecast blocks in the caller's buffer were filled from
    int host_obs_mode = 0;       // glg_set_host_obs_mode
    cudaStream_t own_stream = nullptr;
    cudaEvent_t order_event = nullptr;  // glg_host_path_after
    long long launches = 0;
    double ctrl[GLG_NCTRL];  // rule-based controller settings (defaults: configs/agents/rule_based.yml)
    int obs_nmod = 0, obs_mod[GLG_MAXOBSMOD] = {}, obs_off[GLG_MAXOBSMOD] = {}, fc_off = -1;
    void *nccl_comm = nullptr;  // ncclComm_t (glg_nccl_init)
    // device rollout (glg_rollout_create)
    glg_rollout_config roll = {};
    bool have_roll = false;
    int roll_blocks = 0;
    float *roll_obs = nullptr, *roll_rew = nullptr, *roll_starts = nullptr, *roll_adv = nullptr, *roll_ret = nullptr;
    double *roll_stat = nullptr, *roll_partial = nullptr, *roll_norm = nullptr, *roll_acc = nullptr;
    std::string err;
};
static const double kDefaultCtrl[GLG_NCTRL] = {0, 18, -1, 366, 400, 10, 19.5, 16.5, 0, 5, 800, 4, 85, 2, 5, 1, -1, 5, 10, -1, 4, -2,
                                               2, 2, 100, 85, -1, -100, 1};

#define GLG_CUDA(h, call)                                                                            \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            (h)->err = buf__;                                                                        \
            return GLG_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

static int fail(glg_handle *h, int code, const char *msg) {
    h->err = msg;
    return code;
}

extern "C" void glg_default_config(glg_config *c) {
    memset(c, 0, sizeof *c);
    c->num_envs = 1;
    c->device = 0;
    c->dt = 900.0;
    c->n_sub = 600;
    c->N = 5760;
    c->Np = 48;
    c->precision = 0;
    c->auto_reset = 1;
    for (int i = 0; i < 6; ++i) {
        c->u_min[i] = 0.0;
        c->u_max[i] = 1.0;
    }
    c->delta_u_max = (double)0.1f;
    c->con_low[0] = 300.; c->con_low[1] = 15.; c->con_low[2] = 50.;
    c->con_high[0] = 1600.; c->con_high[1] = 34.; c->con_high[2] = 85.;
    c->elec_price = 0.3; c->heating_price = 0.09; c->co2_price = 0.3; c->fruit_price = 1.6; c->dmfm = 0.065;
    c->fixed_costs = (15. + 0.015 + 0.07 * 116 + 2.) / 365 / (86400 / 900);
    c->uncertainty_scale = 0.0;
    c->seed = 0;
    c->env_id_offset = 0;
    c->role_warps = 0;
    c->role_lanes = 0;
    for (int i = 0; i < 8; ++i) c->obs_modules[i] = 0;  // empty = the default stack of TomatoEnv.yml
}

extern "C" const char *glg_last_error(const glg_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static size_t res_block_bytes(size_t B) { return 17 * B; }  // double reward[B] | int32 timestep[B] | int32 table[B] | uint8 done[B]
template <typename T>
static cudaError_t dev_alloc(T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
}

// ---- episode-statistics all-reduce over NCCL, bound at run time ---------------------------------------------
struct NcclId {  // ncclUniqueId: 128 opaque bytes, passed by value
    char internal[128];
};
namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi *nccl_api(std::string *err) {
    static NcclApi api;
    if (api.lib) return &api;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        *err = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
        return nullptr;
    }
    api.GetUniqueId = (int (*)(void *))dlsym(lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(void **, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t))dlsym(lib, "ncclAllReduce");
    api.CommDestroy = (int (*)(void *))dlsym(lib, "ncclCommDestroy");
    api.GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) {
        *err = "libnccl.so.2 lacks a required symbol";
        return nullptr;
    }
    api.lib = lib;
    return &api;
}

extern "C" void glg_destroy(glg_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->x); cudaFree(h->u); cudaFree(h->time); cudaFree(h->ep_return); cudaFree(h->ep_info);
    cudaFree(h->res_block); cudaFree(h->ep_len); cudaFree(h->step_ctr);
    cudaFree(h->obs_head);
    cudaFree(h->obs); cudaFree(h->term_obs); cudaFree(h->actions); cudaFree(h->info);
    cudaFree(h->stats); cudaFree(h->weather); cudaFree(h->start_day); cudaFree(h->reset_tables);
    if (h->h_actions) cudaFreeHost(h->h_actions);
    if (h->h_obs) cudaFreeHost(h->h_obs);
    if (h->h_reward) cudaFreeHost(h->h_reward);
    if (h->h_done) cudaFreeHost(h->h_done);
    if (h->h_head) cudaFreeHost(h->h_head);
    if (h->h_res) cudaFreeHost(h->h_res);
    if (h->order_event) cudaEventDestroy(h->order_event);
    cudaFree(h->roll_obs); cudaFree(h->roll_rew); cudaFree(h->roll_starts); cudaFree(h->roll_adv); cudaFree(h->roll_ret);
    cudaFree(h->roll_stat); cudaFree(h->roll_partial); cudaFree(h->roll_norm); cudaFree(h->roll_acc);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->nccl_comm) {
        std::string e;
        NcclApi *n = nccl_api(&e);
        if (n) n->CommDestroy(h->nccl_comm);
    }
    delete h;
}

extern "C" int glg_create(const glg_config *cfg, glg_handle **out) {
    if (!cfg || !out) {
        g_create_error = "glg_create: null argument";
        return GLG_ERR_ARG;
    }
    *out = nullptr;
    if (cfg->num_envs < 1 || cfg->n_sub < 1 || cfg->N < 0 || cfg->Np < 0 || !(cfg->dt > 0) ||
        (cfg->precision != 0 && cfg->precision != 1) ||
        (cfg->role_warps < 0 || cfg->role_warps > 3) ||
        (cfg->integrator != 0 && cfg->integrator != 1) || (cfg->precision == 1 && cfg->role_warps == 1)) {
        g_create_error = (cfg->precision == 1 && cfg->role_warps == 1)
                             ? "glg_create: the fp32 throughput mode runs on kernel C only (role_warps 0, 2 or 3)"
                             : "glg_create: invalid num_envs / n_sub / N / Np / dt / precision / role_warps / integrator";
        return GLG_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
        g_create_error = std::string("glg_create: no usable CUDA device (") + cudaGetErrorString(ce) +
                         "); this library has no CPU fallback";
        return GLG_ERR_CUDA;
    }
    glg_handle *h = new (std::nothrow) glg_handle();
    if (!h) return GLG_ERR_ALLOC;
    h->cfg = *cfg;
    for (int i = 0; i < GLG_NCTRL; ++i) h->ctrl[i] = kDefaultCtrl[i];
    h->B = cfg->num_envs;
    {
        // observation row layout from the module list (tomato_env.py:77-96)
        static const int kDefault[6] = {GLG_OBS_CLIMATE, GLG_OBS_CROP, GLG_OBS_CONTROL, GLG_OBS_WEATHER, GLG_OBS_TIME, GLG_OBS_FORECAST};
        const int sizes[8] = {0, GLG_NOBS_STATE, 4, 3, 6, 5, 5, 5 * cfg->Np};
        unsigned seen = 0;
        int off = 0;
        bool ok = true;
        const bool use_default = cfg->obs_modules[0] == 0;
        for (int m = 0; m < GLG_MAXOBSMOD; ++m) {
            const int id = use_default ? (m < 6 ? kDefault[m] : 0) : cfg->obs_modules[m];
            if (id == 0) break;
            if (id < 1 || id > 7 || (seen >> id & 1u)) {
                ok = false;
                break;
            }
            seen |= 1u << id;
            h->obs_mod[h->obs_nmod] = id;
            h->obs_off[h->obs_nmod] = off;
            if (id == GLG_OBS_FORECAST) h->fc_off = off;
            off += sizes[id];
            ++h->obs_nmod;
        }
        if (!ok || off < 3) {
            g_create_error = "glg_create: obs_modules must list distinct module ids 1..7 with at least 3 observation entries in total";
            delete h;
            return GLG_ERR_ARG;
        }
        h->obs_dim = off;
    }
    h->nt = 64;
    // kernel B: envs per CTA.  32 (full warps) is fastest at every batch size measured: a CTA's step time grows with the
    // number of co-resident CTAs faster than under-filled warps could win back (profiles/r1_lane_sweep.txt).
    h->role_lanes = (cfg->role_lanes >= 1 && cfg->role_lanes <= 32) ? cfg->role_lanes : 32;
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, cfg->device) == cudaSuccess) h->sms = prop.multiProcessorCount;
    }
    const size_t B = (size_t)h->B;
    cudaError_t e = cudaSetDevice(cfg->device);
    if (e == cudaSuccess) e = dev_alloc(&h->x, GLG_NX * B);
    if (e == cudaSuccess) e = dev_alloc(&h->u, GLG_NU * B);
    if (e == cudaSuccess) e = dev_alloc(&h->time, 2 * B);
    if (e == cudaSuccess) e = dev_alloc(&h->ep_return, B);
    if (e == cudaSuccess) e = dev_alloc(&h->ep_info, GLG_NINFO * B);
    // reward | timestep | table | done share one allocation: glg_step_host fetches them with one device->host copy
    if (e == cudaSuccess) e = dev_alloc(&h->res_block, res_block_bytes(B));
    if (e == cudaSuccess) {
        h->reward = reinterpret_cast<double *>(h->res_block);
        h->timestep = reinterpret_cast<int *>(h->res_block + 8 * B);
        h->table = h->timestep + B;
        h->done = h->res_block + 16 * B;
    }
    if (e == cudaSuccess) e = dev_alloc(&h->ep_len, B);
    if (e == cudaSuccess) e = dev_alloc(&h->step_ctr, B);
    if (e == cudaSuccess) e = dev_alloc(&h->obs, (size_t)h->obs_dim * B);
    if (e == cudaSuccess) e = dev_alloc(&h->term_obs, (size_t)h->obs_dim * B);
    if (e == cudaSuccess) e = dev_alloc(&h->obs_head, (size_t)(h->obs_dim - (h->fc_off >= 0 ? 5 * cfg->Np : 0)) * B);
    if (e == cudaSuccess) e = dev_alloc(&h->actions, GLG_NU * B);
    if (e == cudaSuccess) e = dev_alloc(&h->info, GLG_NINFO * B);
    if (e == cudaSuccess) e = dev_alloc(&h->stats, (size_t)GLG_NSTATS);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        g_create_error = std::string("glg_create: ") + cudaGetErrorString(e);
        glg_destroy(h);
        return GLG_ERR_CUDA;
    }
    *out = h;
    return GLG_OK;
}

extern "C" int glg_set_seed(glg_handle *h, uint64_t seed) {
    if (!h) return GLG_ERR_ARG;
    h->cfg.seed = seed;
    return GLG_OK;
}

extern "C" int glg_set_params(glg_handle *h, const double *p_host) {
    if (!h || !p_host) return GLG_ERR_ARG;
    for (int i = 0; i < GLG_NP; ++i) h->uni.P[i] = p_host[i];
    glg_make_k(p_host, h->uni.K);
    glg_make_c(p_host, h->uni.C);
    for (int i = 0; i < K_COUNT; ++i) h->uni.Kf[i] = (float)h->uni.K[i];
    for (int i = 0; i < C_COUNT; ++i) h->uni.Cf[i] = (float)h->uni.C[i];
    h->general = !glg_params_nominal_structure(p_host);
    h->have_params = true;
    return GLG_OK;
}

extern "C" int glg_set_weather(glg_handle *h, const double *tables_host, int32_t n_tables, int32_t rows,
                               const double *start_day_host) {
    if (!h || !tables_host || n_tables < 1) return GLG_ERR_ARG;
    if (rows < h->cfg.N + h->cfg.Np + 1) return fail(h, GLG_ERR_ARG, "glg_set_weather: rows < N + Np + 1");
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaFree(h->weather); cudaFree(h->start_day); cudaFree(h->reset_tables);
    h->weather = nullptr; h->start_day = nullptr; h->reset_tables = nullptr;
    const size_t n = (size_t)n_tables * rows * GLG_ND;
    GLG_CUDA(h, cudaMalloc((void **)&h->weather, n * sizeof(double)));
    GLG_CUDA(h, cudaMemcpy(h->weather, tables_host, n * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<double> sd(n_tables, 0.0);
    if (start_day_host) sd.assign(start_day_host, start_day_host + n_tables);
    GLG_CUDA(h, cudaMalloc((void **)&h->start_day, n_tables * sizeof(double)));
    GLG_CUDA(h, cudaMemcpy(h->start_day, sd.data(), n_tables * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<int> ids(n_tables);
    for (int i = 0; i < n_tables; ++i) ids[i] = i;
    GLG_CUDA(h, cudaMalloc((void **)&h->reset_tables, n_tables * sizeof(int)));
    GLG_CUDA(h, cudaMemcpy(h->reset_tables, ids.data(), n_tables * sizeof(int), cudaMemcpyHostToDevice));
    h->n_tables = n_tables;
    h->rows = rows;
    h->n_reset_tables = n_tables;
    h->have_weather = true;
    // host copy of what glg_write_forecast shows of the bank (columns 0..4, double -> float), for glg_step_host
    h->fc_bank.resize((size_t)n_tables * rows * 5);
    for (size_t r = 0; r < (size_t)n_tables * rows; ++r)
        for (int c = 0; c < 5; ++c) h->fc_bank[r * 5 + c] = (float)tables_host[r * GLG_ND + c];
    h->kt_valid = false;
    return GLG_OK;
}

extern "C" int glg_set_reset_tables(glg_handle *h, const int32_t *ids, int32_t n) {
    if (!h || !ids || n < 1) return GLG_ERR_ARG;
    if (!h->have_weather) return fail(h, GLG_ERR_STATE, "glg_set_reset_tables: set the weather bank first");
    for (int i = 0; i < n; ++i)
        if (ids[i] < 0 || ids[i] >= h->n_tables) return fail(h, GLG_ERR_ARG, "glg_set_reset_tables: id out of range");
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaFree(h->reset_tables);
    h->reset_tables = nullptr;
    GLG_CUDA(h, cudaMalloc((void **)&h->reset_tables, n * sizeof(int)));
    GLG_CUDA(h, cudaMemcpy(h->reset_tables, ids, n * sizeof(int), cudaMemcpyHostToDevice));
    h->n_reset_tables = n;
    return GLG_OK;
}

static void fill_args(const glg_handle *h, GlgStepArgs *a) {
    memset(a, 0, sizeof *a);
    const glg_config &c = h->cfg;
    a->B = h->B; a->n_sub = c.n_sub; a->N = c.N; a->Np = c.Np; a->rows = h->rows; a->n_tables = h->n_tables;
    a->obs_dim = h->obs_dim; a->auto_reset = c.auto_reset; a->n_reset_tables = h->n_reset_tables;
    a->dt = c.dt;
    for (int i = 0; i < GLG_NU; ++i) {
        a->u_min[i] = c.u_min[i];
        a->u_max[i] = c.u_max[i];
    }
    a->delta_u_max_f32 = (float)c.delta_u_max;
    for (int i = 0; i < 3; ++i) {
        a->con_low[i] = c.con_low[i];
        a->con_high[i] = c.con_high[i];
    }
    a->elec_price = c.elec_price; a->heating_price = c.heating_price; a->co2_price = c.co2_price;
    a->fruit_price = c.fruit_price; a->dmfm = c.dmfm; a->fixed_costs = c.fixed_costs;
    a->uncertainty_scale = c.uncertainty_scale;
    a->seed = c.seed; a->env_id_offset = c.env_id_offset;
    a->role_lanes = h->role_lanes;
    a->integrator = c.integrator;
    a->obs_nmod = h->obs_nmod;
    for (int i = 0; i < GLG_MAXOBSMOD; ++i) {
        a->obs_mod[i] = h->obs_mod[i];
        a->obs_off[i] = h->obs_off[i];
    }
    a->fc_off = h->fc_off;
    a->weather = h->weather; a->start_day = h->start_day; a->reset_tables = h->reset_tables;
    a->x = h->x; a->u = h->u; a->time = h->time; a->ep_return = h->ep_return; a->ep_info = h->ep_info;
    a->timestep = h->timestep; a->table = h->table; a->ep_len = h->ep_len; a->step_ctr = h->step_ctr;
    a->obs_head = h->obs_head;
    a->obs = h->obs; a->term_obs = h->term_obs; a->reward = h->reward; a->info = h->info; a->stats = h->stats;
    a->done = h->done;
}

static int check_ready(glg_handle *h) {
    if (!h) return GLG_ERR_ARG;
    if (!h->have_params) return fail(h, GLG_ERR_STATE, "parameters not set (glg_set_params)");
    if (!h->have_weather) return fail(h, GLG_ERR_STATE, "weather bank not set (glg_set_weather)");
    return GLG_OK;
}

extern "C" int glg_reset(glg_handle *h, const uint8_t *mask_dev, const int32_t *table_ids_dev, void *stream) {
    int rc = check_ready(h);
    if (rc) return rc;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    GlgStepArgs a;
    fill_args(h, &a);
    const int NT = 128;
    glg_reset_kernel<NT><<<(h->B + NT - 1) / NT, NT, 0, (cudaStream_t)stream>>>(a, mask_dev, table_ids_dev);
    h->launches += 1;
    GLG_CUDA(h, cudaGetLastError());
    h->is_reset = true;
    h->kt_valid = false;
    return GLG_OK;
}

template <bool GENERAL, bool NOISY>
static cudaError_t launch_step(glg_handle *h, const GlgStepArgs &a, cudaStream_t s) {
    constexpr int NT = 64;
    const size_t smem = GlgStepSmem<NT, NOISY>::bytes(a.Np);
    static size_t attr_smem_dev[64] = {};  // largest size opted into so far, per device (the attribute is per context)
    size_t &attr_smem = attr_smem_dev[h->cfg.device & 63];
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(glg_step_kernel<GENERAL, NOISY, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem = smem;
    }
    glg_step_kernel<GENERAL, NOISY, NT><<<(a.B + NT - 1) / NT, NT, smem, s>>>(h->uni, a);
    return cudaGetLastError();
}

// NG group warps for the fp64 / fp32 unit kernels; MINB = CTAs per SM the register budget is compiled for
#ifndef GLG_NG
#define GLG_NG 12
#endif
#ifndef GLG_MINB_LAT
#define GLG_MINB_LAT 1   // CTAs per SM of the latency variant (role_warps = 2)
#endif
#ifndef GLG_MINB_TPUT
#define GLG_MINB_TPUT 4  // CTAs per SM of the throughput variant (role_warps = 3)
#endif
#ifndef GLG_TPUT_FUSED
#define GLG_TPUT_FUSED true  // throughput variant: fused layout (4 warps = group role + owner), else dedicated owner warps
#endif
template <class T, bool GENERAL, bool NOISY, int MINB, bool FUSED>
static cudaError_t launch_step_units_t(glg_handle *h, const GlgStepArgs &a, cudaStream_t s) {
    constexpr int NG = FUSED ? GLG_NO : GLG_NG;
    const size_t smem = GlgRoleSmem<T, NG, GENERAL, NOISY>::bytes(a.Np);
    static size_t attr_smem_dev[64] = {};  // largest size opted into so far, per device (the attribute is per context)
    size_t &attr_smem = attr_smem_dev[h->cfg.device & 63];
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(glg_step_units_kernel<T, GENERAL, NOISY, NG, MINB, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_smem = smem;
    }
    glg_step_units_kernel<T, GENERAL, NOISY, NG, MINB, FUSED><<<(a.B + a.role_lanes - 1) / a.role_lanes, 32 * (NG + (FUSED ? 0 : GLG_NO)), smem, s>>>(h->uni, a);
    return cudaGetLastError();
}
template <bool GENERAL, bool NOISY, int MINB, bool FUSED>
static cudaError_t launch_step_units(glg_handle *h, const GlgStepArgs &a, cudaStream_t s) {
    // precision 0: fp64 parity mode ; 1: flux units in fp32, RK4 state / stage sums / epilogue in fp64
#ifdef GLG_DEV_FAST  // experimental builds (tools/ubench): only the fp64 nominal variants are compiled
    if (GENERAL || NOISY || h->cfg.precision == 1) return cudaErrorNotSupported;
    return launch_step_units_t<double, false, false, MINB, FUSED>(h, a, s);
#else
    return h->cfg.precision == 1 ? launch_step_units_t<float, GENERAL, NOISY, MINB, FUSED>(h, a, s)
                                 : launch_step_units_t<double, GENERAL, NOISY, MINB, FUSED>(h, a, s);
#endif
}

// Kernel variant: 1 = kernel A (one thread per env), 2 = kernel C latency layout (4 owner + 12 group warps per 32 envs, one CTA
// per SM: the step time of a batch that leaves SMs under-filled is the latency of one CTA), 3 = kernel C throughput layout
// (4 fused warps per 32 envs, 4 CTAs per SM).  Auto-pick: measured with tools/layout_sweep.py the latency layout wins up to two
// waves of CTAs (B = 9472 on 148 SMs: 1.735 vs 1.814 ms), the throughput layout beyond (B = 12 288: 2.60 vs 2.18 ms).
static int pick_role_warps(const glg_handle *h) {
    if (h->cfg.role_warps != 0) return h->cfg.role_warps;
    return h->B <= 2 * h->sms * GLG_ROLE_LANES ? 2 : 3;
}

static int step_common(glg_handle *h, const float *actions_dev, const double *controls_dev, const double *noise_dev,
                       void *stream, bool rule_based = false) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!h->is_reset) return fail(h, GLG_ERR_STATE, "step before reset");
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    h->kt_valid = false;
    GlgStepArgs a;
    fill_args(h, &a);
    a.actions = actions_dev;
    a.controls = controls_dev;
    a.raw_control = rule_based ? 2 : (controls_dev ? 1 : 0);
    for (int i = 0; i < GLG_NCTRL; ++i) a.ctrl[i] = h->ctrl[i];
    a.noise = noise_dev;
    const bool noisy = (h->cfg.uncertainty_scale != 0.0) || (noise_dev != nullptr);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    const int rw = pick_role_warps(h);
    if (rw == 2) {
        if (h->general) e = noisy ? launch_step_units<true, true, GLG_MINB_LAT, false>(h, a, s) : launch_step_units<true, false, GLG_MINB_LAT, false>(h, a, s);
        else e = noisy ? launch_step_units<false, true, GLG_MINB_LAT, false>(h, a, s) : launch_step_units<false, false, GLG_MINB_LAT, false>(h, a, s);
    } else if (rw == 3) {
        if (h->general) e = noisy ? launch_step_units<true, true, GLG_MINB_TPUT, GLG_TPUT_FUSED>(h, a, s) : launch_step_units<true, false, GLG_MINB_TPUT, GLG_TPUT_FUSED>(h, a, s);
        else e = noisy ? launch_step_units<false, true, GLG_MINB_TPUT, GLG_TPUT_FUSED>(h, a, s) : launch_step_units<false, false, GLG_MINB_TPUT, GLG_TPUT_FUSED>(h, a, s);
    } else {
        if (h->general) e = noisy ? launch_step<true, true>(h, a, s) : launch_step<true, false>(h, a, s);
        else e = noisy ? launch_step<false, true>(h, a, s) : launch_step<false, false>(h, a, s);
    }
    h->launches += 1;
    GLG_CUDA(h, e);
    return GLG_OK;
}

extern "C" int glg_step(glg_handle *h, const float *actions_dev, const double *noise_dev, void *stream) {
    if (!h || !actions_dev) return GLG_ERR_ARG;
    return step_common(h, actions_dev, nullptr, noise_dev, stream);
}

extern "C" int glg_step_raw_control(glg_handle *h, const double *controls_dev, const double *noise_dev, void *stream) {
    if (!h || !controls_dev) return GLG_ERR_ARG;
    return step_common(h, nullptr, controls_dev, noise_dev, stream);
}

extern "C" int glg_set_rule_controller(glg_handle *h, const double *settings29) {
    if (!h) return GLG_ERR_ARG;
    for (int i = 0; i < GLG_NCTRL; ++i) h->ctrl[i] = settings29 ? settings29[i] : kDefaultCtrl[i];
    return GLG_OK;
}

extern "C" int glg_step_rule_based(glg_handle *h, const double *noise_dev, void *stream) {
    if (!h) return GLG_ERR_ARG;
    return step_common(h, nullptr, nullptr, noise_dev, stream, true);
}

// known-answer entry for the controller alone: one thread per point
__global__ void glg_rule_control_kernel(GlgStepArgs A, const double *x, const double *d, const double *hod, const double *doy,
                                        double *u, int n) {
    glg_exp_tbl_fill();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xl[GLG_NX], dl[GLG_ND], ul[GLG_NU];
    for (int j = 0; j < GLG_NX; ++j) xl[j] = x[(size_t)i * GLG_NX + j];
    for (int j = 0; j < GLG_ND; ++j) dl[j] = d[(size_t)i * GLG_ND + j];
    glg_rule_control(A.ctrl, xl, dl, hod[i], doy[i], ul);
    for (int j = 0; j < GLG_NU; ++j) u[(size_t)i * GLG_NU + j] = ul[j];
}
extern "C" int glg_rule_control_batch(const double *settings29, const double *x_dev, const double *d_dev, const double *hod_dev,
                                      const double *doy_dev, double *u_dev, int32_t n, int32_t device, void *stream) {
    if (!x_dev || !d_dev || !hod_dev || !doy_dev || !u_dev || n < 1) {
        g_create_error = "glg_rule_control_batch: invalid argument";
        return GLG_ERR_ARG;
    }
    GlgStepArgs a;
    memset(&a, 0, sizeof a);
    for (int i = 0; i < GLG_NCTRL; ++i) a.ctrl[i] = settings29 ? settings29[i] : kDefaultCtrl[i];
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        glg_rule_control_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a, x_dev, d_dev, hod_dev, doy_dev, u_dev, n);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) {
        g_create_error = std::string("glg_rule_control_batch: ") + cudaGetErrorString(e);
        return GLG_ERR_CUDA;
    }
    return GLG_OK;
}

static bool is_pinned(const void *p) {
    // small per-thread cache: the callers pass the same few buffers every step, and a stale answer is harmless (a pageable
    // buffer taken for page-locked makes cudaMemcpyAsync stage it itself; the reverse costs one staging copy)
    static thread_local struct { const void *p; bool pinned; } cache[8] = {};
    static thread_local int next = 0;
    for (auto &c : cache)
        if (c.p == p) return c.pinned;
    cudaPointerAttributes at;
    bool pinned = false;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) cudaGetLastError();
    else pinned = at.type == cudaMemoryTypeHost;
    cache[next] = {p, pinned};
    next = (next + 1) % 8;
    return pinned;
}

// Rows [lo, hi) of a per-env host loop on up to 8 threads; small batches run inline (a thread costs ~20 us to start).  A chunk
// whose thread cannot be started (no exception may cross the C boundary) is done by the caller.
template <class F>
static void host_rows(size_t n_rows, size_t bytes_per_row, const F &f) {
    const size_t total = n_rows * bytes_per_row;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = total / ((size_t)6 << 20);  // one thread per 6 MB moved
    const size_t cap = hw >= 4 ? (hw / 2 < 8 ? hw / 2 : 8) : 1;
    nt = nt > cap ? cap : nt;
    if (nt <= 1) {
        f((size_t)0, n_rows);
        return;
    }
    std::vector<std::thread> th;
    const size_t chunk = (n_rows + nt - 1) / nt;
    for (size_t t = 1; t < nt; ++t) {
        const size_t lo = t * chunk, hi = lo + chunk < n_rows ? lo + chunk : n_rows;
        if (lo >= hi) continue;
        try {
            th.emplace_back([&f, lo, hi] { f(lo, hi); });
        } catch (...) {
            f(lo, hi);
        }
    }
    f((size_t)0, chunk < n_rows ? chunk : n_rows);
    for (auto &t : th) t.join();
}

extern "C" int glg_host_path_after(glg_handle *h, void *stream) {
    if (!h) return GLG_ERR_ARG;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    if (!h->order_event) GLG_CUDA(h, cudaEventCreateWithFlags(&h->order_event, cudaEventDisableTiming));
    GLG_CUDA(h, cudaEventRecord(h->order_event, (cudaStream_t)stream));
    GLG_CUDA(h, cudaStreamWaitEvent(h->own_stream, h->order_event, 0));
    return GLG_OK;
}

extern "C" int glg_set_host_obs_mode(glg_handle *h, int32_t mode) {
    if (!h || (mode != 0 && mode != 1)) return GLG_ERR_ARG;
    h->host_obs_mode = mode;
    return GLG_OK;
}

// Overlapped observation path (host_obs_mode 0, the default, for stacks with a WeatherForecastObservations block): 5 Np of the
// row's floats (240 of 263 in the default stack) are weather rows kw+1 .. kw+Np of the env's table, kw = min(timestep before the
// step, rows - Np - 1) -- known before the kernel runs.  So only the rest of the row (glg_write_obs_row's packed copy, obs_head)
// crosses PCIe, and WHILE THE KERNEL RUNS the host writes the forecast blocks into the caller's buffer from its float32 copy of
// the bank, predicted from the (timestep, table) of the previous host step.  After the synchronise the packed columns are
// scattered into the rows and every prediction is checked against the (timestep, table) the step left behind: rows that reset
// in place (auto-reset draws a new table on the device), and all rows when something other than glg_step_host moved the envs in
// between (reset, tensor steps, set_state), are filled again from the verified values.  The result is the same [B][obs_dim]
// array the full device->host copy (mode 1) delivers, bit for bit (tests/test_gpu_features.py).
static int step_host_overlapped(glg_handle *h, const float *a_src, float *obs_host, double *reward_host, uint8_t *done_host, bool o_pin) {
    const size_t B = (size_t)h->B;
    const int Np = h->cfg.Np, nf = 5 * Np, D = h->obs_dim, n_head = D - nf, fc = h->fc_off, rows = h->rows;
    if (!h->h_res) {
        GLG_CUDA(h, cudaMallocHost((void **)&h->h_res, 2 * res_block_bytes(B)));
        h->fc_pred.assign(2 * B, -1);
        h->kt_valid = false;
    }
    if (!o_pin && !h->h_head) GLG_CUDA(h, cudaMallocHost((void **)&h->h_head, (size_t)n_head * B * sizeof(float)));
    cudaStream_t s = h->own_stream;
    // result block of the previous host step (prediction source) / of this one
    const unsigned char *res_cur = h->h_res + (size_t)h->kt_cur * res_block_bytes(B);
    unsigned char *res_nxt = h->h_res + (size_t)(h->kt_cur ^ 1) * res_block_bytes(B);
    const int *cur = reinterpret_cast<const int *>(res_cur + 8 * B);
    const int *nxt = reinterpret_cast<const int *>(res_nxt + 8 * B);
    const bool valid = h->kt_valid;  // every entry that moves the envs clears it (step_common included: set again below)
    GLG_CUDA(h, cudaMemcpyAsync(h->actions, a_src, GLG_NU * B * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = step_common(h, h->actions, nullptr, nullptr, s);
    if (rc) return rc;
    if (o_pin) {
        // page-locked destination: the copy engine puts the packed columns straight into the rows (strided 2-D copy)
        if (fc > 0)
            GLG_CUDA(h, cudaMemcpy2DAsync(obs_host, (size_t)D * sizeof(float), h->obs_head, (size_t)n_head * sizeof(float), (size_t)fc * sizeof(float), B,
                                          cudaMemcpyDeviceToHost, s));
        if (n_head > fc)
            GLG_CUDA(h, cudaMemcpy2DAsync(obs_host + fc + nf, (size_t)D * sizeof(float), h->obs_head + fc, (size_t)n_head * sizeof(float),
                                          (size_t)(n_head - fc) * sizeof(float), B, cudaMemcpyDeviceToHost, s));
    } else {
        GLG_CUDA(h, cudaMemcpyAsync(h->h_head, h->obs_head, (size_t)n_head * B * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    GLG_CUDA(h, cudaMemcpyAsync(res_nxt, h->res_block, res_block_bytes(B), cudaMemcpyDeviceToHost, s));  // reward | timestep | table | done
    // ---- while the GPU works: forecast blocks from the predicted (table, row)
    const float *bank = h->fc_bank.data();
    int *pred_t = h->fc_pred.data(), *pred_r = pred_t + B;
    const int n_tables = h->n_tables, kmax = rows - Np - 1;
    host_rows(B, (size_t)nf * sizeof(float), [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const int k = valid ? cur[i] : -1, t = valid ? cur[B + i] : -1;
            if (k < 0 || t < 0 || t >= n_tables) {
                pred_t[i] = -1;
                continue;
            }
            const int r0 = (k < kmax ? k : kmax) + 1;
            memcpy(obs_host + i * D + fc, bank + ((size_t)t * rows + r0) * 5, (size_t)nf * sizeof(float));
            pred_t[i] = t;
            pred_r[i] = r0;
        }
    });
    GLG_CUDA(h, cudaStreamSynchronize(s));
    // ---- verify every prediction against what the step left behind (and, for a pageable destination, scatter the packed columns)
    const float *head = h->h_head;
    host_rows(B, o_pin ? 16 : (size_t)D * sizeof(float) / 4, [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            float *row = obs_host + i * D;
            if (!o_pin) {
                const float *hd = head + i * n_head;
                if (fc > 0) memcpy(row, hd, (size_t)fc * sizeof(float));
                if (n_head > fc) memcpy(row + fc + nf, hd + fc, (size_t)(n_head - fc) * sizeof(float));
            }
            // timestep after the step: 0 = reset in place (row 1 of the new table), else the pre-step timestep + 1
            const int kp = nxt[i], t = nxt[B + i];
            const int r0 = kp <= 0 ? 1 : ((kp - 1 < kmax ? kp - 1 : kmax) + 1);
            if (t != pred_t[i] || r0 != pred_r[i])
                memcpy(row + fc, bank + ((size_t)t * rows + r0) * 5, (size_t)nf * sizeof(float));
        }
    });
    if (reward_host) memcpy(reward_host, res_nxt, B * sizeof(double));
    if (done_host) memcpy(done_host, res_nxt + 16 * B, B);
    h->kt_cur ^= 1;
    h->kt_valid = true;
    return GLG_OK;
}

extern "C" int glg_step_host(glg_handle *h, const float *actions_host, float *obs_host, double *reward_host,
                             uint8_t *done_host) {
    if (!h || !actions_host) return GLG_ERR_ARG;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t B = (size_t)h->B;
    if (!h->h_actions) {
        GLG_CUDA(h, cudaMallocHost((void **)&h->h_actions, GLG_NU * B * sizeof(float)));
        GLG_CUDA(h, cudaMallocHost((void **)&h->h_reward, B * sizeof(double)));
        GLG_CUDA(h, cudaMallocHost((void **)&h->h_done, B));
    }
    cudaStream_t s = h->own_stream;
    // caller buffers that are page-locked are used directly; pageable ones are staged through the handle's pinned buffers
    const bool a_pin = is_pinned(actions_host);
    const bool o_pin = obs_host && is_pinned(obs_host), r_pin = reward_host && is_pinned(reward_host),
               d_pin = done_host && is_pinned(done_host);
    const float *a_src = actions_host;
    if (!a_pin) {
        memcpy(h->h_actions, actions_host, GLG_NU * B * sizeof(float));
        a_src = h->h_actions;
    }
    if (obs_host && h->host_obs_mode == 0 && h->fc_off >= 0 && !h->fc_bank.empty()) {
        return step_host_overlapped(h, a_src, obs_host, reward_host, done_host, o_pin);
    } else {
        h->kt_valid = false;
        if (obs_host && !o_pin && !h->h_obs) GLG_CUDA(h, cudaMallocHost((void **)&h->h_obs, (size_t)h->obs_dim * B * sizeof(float)));
        GLG_CUDA(h, cudaMemcpyAsync(h->actions, a_src, GLG_NU * B * sizeof(float), cudaMemcpyHostToDevice, s));
        int rc = step_common(h, h->actions, nullptr, nullptr, s);
        if (rc) return rc;
        if (obs_host)
            GLG_CUDA(h, cudaMemcpyAsync(o_pin ? obs_host : h->h_obs, h->obs, (size_t)h->obs_dim * B * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (reward_host) GLG_CUDA(h, cudaMemcpyAsync(r_pin ? reward_host : h->h_reward, h->reward, B * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (done_host) GLG_CUDA(h, cudaMemcpyAsync(d_pin ? done_host : h->h_done, h->done, B, cudaMemcpyDeviceToHost, s));
        GLG_CUDA(h, cudaStreamSynchronize(s));
        if (obs_host && !o_pin) memcpy(obs_host, h->h_obs, (size_t)h->obs_dim * B * sizeof(float));
    }
    if (reward_host && !r_pin) memcpy(reward_host, h->h_reward, B * sizeof(double));
    if (done_host && !d_pin) memcpy(done_host, h->h_done, B);
    return GLG_OK;
}

extern "C" int glg_step_host_split(glg_handle *h, const float *actions_host, float *head_host, int32_t *timestep_host,
                                   int32_t *table_host, double *reward_host, uint8_t *done_host) {
    if (!h || !actions_host) return GLG_ERR_ARG;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t B = (size_t)h->B;
    cudaStream_t s = h->own_stream;
    // pageable buffers work too (the runtime stages them); page-locked ones (the Python wrapper's) make the copies asynchronous
    GLG_CUDA(h, cudaMemcpyAsync(h->actions, actions_host, GLG_NU * B * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = step_common(h, h->actions, nullptr, nullptr, s);
    if (rc) return rc;
    if (head_host) {  // the kernels write the row without its forecast block a second time, packed: one contiguous copy
        const size_t n_head = (size_t)(h->obs_dim - (h->fc_off >= 0 ? 5 * h->cfg.Np : 0));
        GLG_CUDA(h, cudaMemcpyAsync(head_host, h->obs_head, n_head * B * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    if (timestep_host) GLG_CUDA(h, cudaMemcpyAsync(timestep_host, h->timestep, B * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (table_host) GLG_CUDA(h, cudaMemcpyAsync(table_host, h->table, B * sizeof(int), cudaMemcpyDeviceToHost, s));
    if (reward_host) GLG_CUDA(h, cudaMemcpyAsync(reward_host, h->reward, B * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (done_host) GLG_CUDA(h, cudaMemcpyAsync(done_host, h->done, B, cudaMemcpyDeviceToHost, s));
    GLG_CUDA(h, cudaStreamSynchronize(s));
    return GLG_OK;
}

extern "C" int32_t glg_obs_dim(const glg_handle *h) { return h ? h->obs_dim : 0; }
extern "C" float *glg_obs_dev(glg_handle *h) { return h ? h->obs : nullptr; }
extern "C" float *glg_terminal_obs_dev(glg_handle *h) { return h ? h->term_obs : nullptr; }
extern "C" double *glg_reward_dev(glg_handle *h) { return h ? h->reward : nullptr; }
extern "C" uint8_t *glg_done_dev(glg_handle *h) { return h ? h->done : nullptr; }
extern "C" double *glg_info_dev(glg_handle *h) { return h ? h->info : nullptr; }
extern "C" double *glg_state_dev(glg_handle *h) { return h ? h->x : nullptr; }
extern "C" double *glg_controls_dev(glg_handle *h) { return h ? h->u : nullptr; }
extern "C" int32_t *glg_timestep_dev(glg_handle *h) { return h ? h->timestep : nullptr; }
extern "C" int32_t *glg_table_dev(glg_handle *h) { return h ? h->table : nullptr; }
extern "C" double *glg_time_dev(glg_handle *h) { return h ? h->time : nullptr; }
extern "C" double *glg_stats_dev(glg_handle *h) { return h ? h->stats : nullptr; }
extern "C" int64_t glg_launch_count(const glg_handle *h) { return h ? h->launches : 0; }

extern "C" int glg_clear_stats(glg_handle *h, void *stream) {
    if (!h) return GLG_ERR_ARG;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    GLG_CUDA(h, cudaMemsetAsync(h->stats, 0, GLG_NSTATS * sizeof(double), (cudaStream_t)stream));
    return GLG_OK;
}

// [B][n] row-major host  <->  [n][B] device SoA
static void to_soa(const double *aos, double *soa, int B, int n) {
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < n; ++i) soa[(size_t)i * B + b] = aos[(size_t)b * n + i];
}
static void to_aos(const double *soa, double *aos, int B, int n) {
    for (int b = 0; b < B; ++b)
        for (int i = 0; i < n; ++i) aos[(size_t)b * n + i] = soa[(size_t)i * B + b];
}

extern "C" int glg_set_state(glg_handle *h, const double *x_host, const double *u_host, const int32_t *timestep_host) {
    if (!h) return GLG_ERR_ARG;
    h->kt_valid = false;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    GLG_CUDA(h, cudaDeviceSynchronize());
    const int B = h->B;
    std::vector<double> tmp;
    if (x_host) {
        tmp.resize((size_t)GLG_NX * B);
        to_soa(x_host, tmp.data(), B, GLG_NX);
        GLG_CUDA(h, cudaMemcpy(h->x, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (u_host) {
        tmp.resize((size_t)GLG_NU * B);
        to_soa(u_host, tmp.data(), B, GLG_NU);
        GLG_CUDA(h, cudaMemcpy(h->u, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (timestep_host) {
        // the env clock follows the timestep (tomato_env.py:126-128, :246-247): doy = start_day(table) + k dt/86400 (summed the way
        // the step does, one increment per step), hod = (k dt/3600) mod 24
        if (!h->have_weather) return fail(h, GLG_ERR_STATE, "glg_set_state: set the weather bank first");
        for (int b = 0; b < B; ++b)
            if (timestep_host[b] < 0 || timestep_host[b] > h->cfg.N) return fail(h, GLG_ERR_ARG, "glg_set_state: timestep outside [0, N]");
        std::vector<int> tb(B);
        std::vector<double> sd(h->n_tables), tm((size_t)2 * B);
        GLG_CUDA(h, cudaMemcpy(tb.data(), h->table, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost));
        GLG_CUDA(h, cudaMemcpy(sd.data(), h->start_day, (size_t)h->n_tables * sizeof(double), cudaMemcpyDeviceToHost));
        const double ddoy = fmod(h->cfg.dt / 86400.0, 365.0), dhod = h->cfg.dt / 3600.0;
        for (int b = 0; b < B; ++b) {
            double doy = sd[tb[b]], hod = 0.0;
            for (int k = 0; k < timestep_host[b]; ++k) {
                doy += ddoy;
                hod = fmod(hod + dhod, 24.0);
            }
            tm[b] = doy;
            tm[(size_t)B + b] = hod;
        }
        GLG_CUDA(h, cudaMemcpy(h->time, tm.data(), tm.size() * sizeof(double), cudaMemcpyHostToDevice));
        GLG_CUDA(h, cudaMemcpy(h->timestep, timestep_host, (size_t)B * sizeof(int), cudaMemcpyHostToDevice));
    }
    return GLG_OK;
}

extern "C" int glg_get_state(glg_handle *h, double *x_host, double *u_host, int32_t *timestep_host) {
    if (!h) return GLG_ERR_ARG;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    GLG_CUDA(h, cudaDeviceSynchronize());
    const int B = h->B;
    std::vector<double> tmp;
    if (x_host) {
        tmp.resize((size_t)GLG_NX * B);
        GLG_CUDA(h, cudaMemcpy(tmp.data(), h->x, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        to_aos(tmp.data(), x_host, B, GLG_NX);
    }
    if (u_host) {
        tmp.resize((size_t)GLG_NU * B);
        GLG_CUDA(h, cudaMemcpy(tmp.data(), h->u, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        to_aos(tmp.data(), u_host, B, GLG_NU);
    }
    if (timestep_host) GLG_CUDA(h, cudaMemcpy(timestep_host, h->timestep, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost));
    return GLG_OK;
}

// full per-env state: checkpoint / restore
template <class T>
static cudaError_t copy_field(bool to_dev, T *dev, T *host, size_t n_per_env, int B, bool soa) {
    if (!host) return cudaSuccess;
    const size_t n = n_per_env * (size_t)B;
    if (!soa || n_per_env == 1) return to_dev ? cudaMemcpy(dev, host, n * sizeof(T), cudaMemcpyHostToDevice) : cudaMemcpy(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost);
    std::vector<T> tmp(n);
    if (to_dev) {
        for (int b = 0; b < B; ++b)
            for (size_t i = 0; i < n_per_env; ++i) tmp[i * B + b] = host[(size_t)b * n_per_env + i];
        return cudaMemcpy(dev, tmp.data(), n * sizeof(T), cudaMemcpyHostToDevice);
    }
    cudaError_t e = cudaMemcpy(tmp.data(), dev, n * sizeof(T), cudaMemcpyDeviceToHost);
    for (int b = 0; b < B; ++b)
        for (size_t i = 0; i < n_per_env; ++i) host[(size_t)b * n_per_env + i] = tmp[i * B + b];
    return e;
}
static int state_ex(glg_handle *h, const glg_env_state *s, bool to_dev) {
    if (!h || !s) return GLG_ERR_ARG;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    GLG_CUDA(h, cudaDeviceSynchronize());
    const int B = h->B;
    if (to_dev) h->kt_valid = false;
    if (to_dev) {
        if (s->table) {
            if (!h->have_weather) return fail(h, GLG_ERR_STATE, "glg_set_state_ex: set the weather bank first");
            for (int b = 0; b < B; ++b)
                if (s->table[b] < 0 || s->table[b] >= h->n_tables) return fail(h, GLG_ERR_ARG, "glg_set_state_ex: table id out of range");
        }
        if (s->timestep)
            for (int b = 0; b < B; ++b)
                if (s->timestep[b] < 0 || s->timestep[b] > h->cfg.N + 1) return fail(h, GLG_ERR_ARG, "glg_set_state_ex: timestep outside [0, N + 1]");
    }
    GLG_CUDA(h, copy_field(to_dev, h->x, s->x, GLG_NX, B, true));
    GLG_CUDA(h, copy_field(to_dev, h->u, s->u, GLG_NU, B, true));
    GLG_CUDA(h, copy_field(to_dev, h->timestep, s->timestep, 1, B, false));
    GLG_CUDA(h, copy_field(to_dev, h->table, s->table, 1, B, false));
    GLG_CUDA(h, copy_field(to_dev, h->time, s->time, 2, B, true));
    GLG_CUDA(h, copy_field(to_dev, h->step_ctr, s->step_ctr, 1, B, false));
    GLG_CUDA(h, copy_field(to_dev, h->ep_return, s->ep_return, 1, B, false));
    GLG_CUDA(h, copy_field(to_dev, h->ep_len, s->ep_len, 1, B, false));
    GLG_CUDA(h, copy_field(to_dev, h->ep_info, s->ep_info, GLG_NINFO, B, true));
    if (to_dev) h->is_reset = true;
    return GLG_OK;
}
extern "C" int glg_get_state_ex(glg_handle *h, const glg_env_state *out_host) { return state_ex(h, out_host, false); }
extern "C" int glg_set_state_ex(glg_handle *h, const glg_env_state *in_host) { return state_ex(h, in_host, true); }

// ---- device rollout: VecNormalize + rollout buffer + GAE (glg_rollout.cuh) -------------------------------------
extern "C" int glg_rollout_create(glg_handle *h, const glg_rollout_config *cfg) {
    if (!h || !cfg) return GLG_ERR_ARG;
    if (cfg->n_steps < 1 || !(cfg->gamma >= 0.0 && cfg->gamma <= 1.0) || !(cfg->gae_lambda >= 0.0 && cfg->gae_lambda <= 1.0) ||
        !(cfg->clip_obs > 0) || !(cfg->clip_reward > 0) || !(cfg->epsilon > 0))
        return fail(h, GLG_ERR_ARG, "glg_rollout_create: invalid n_steps / gamma / gae_lambda / clip / epsilon");
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaFree(h->roll_obs); cudaFree(h->roll_rew); cudaFree(h->roll_starts); cudaFree(h->roll_adv); cudaFree(h->roll_ret);
    cudaFree(h->roll_stat); cudaFree(h->roll_partial); cudaFree(h->roll_norm); cudaFree(h->roll_acc);
    h->roll_obs = h->roll_rew = h->roll_starts = h->roll_adv = h->roll_ret = nullptr;
    h->roll_stat = h->roll_partial = h->roll_norm = h->roll_acc = nullptr;
    h->have_roll = false;
    const size_t B = (size_t)h->B, T = (size_t)cfg->n_steps, D = (size_t)h->obs_dim;
    h->roll_blocks = (int)((B + GLG_ROLL_ROWS - 1) / GLG_ROLL_ROWS);
    GLG_CUDA(h, dev_alloc(&h->roll_obs, (T + 1) * B * D));
    GLG_CUDA(h, dev_alloc(&h->roll_rew, T * B));
    GLG_CUDA(h, dev_alloc(&h->roll_starts, (T + 1) * B));
    GLG_CUDA(h, dev_alloc(&h->roll_adv, T * B));
    GLG_CUDA(h, dev_alloc(&h->roll_ret, T * B));
    GLG_CUDA(h, dev_alloc(&h->roll_stat, (D + 1) * 3));
    GLG_CUDA(h, dev_alloc(&h->roll_partial, (size_t)h->roll_blocks * (D + 1) * 2));
    GLG_CUDA(h, dev_alloc(&h->roll_norm, (D + 1) * 2));
    GLG_CUDA(h, dev_alloc(&h->roll_acc, B));
    std::vector<double> st((D + 1) * 3);
    for (size_t c = 0; c <= D; ++c) {  // RunningMeanStd: mean 0, var 1, count 1e-4
        st[3 * c] = 0.0;
        st[3 * c + 1] = 1.0;
        st[3 * c + 2] = 1e-4;
    }
    GLG_CUDA(h, cudaMemcpy(h->roll_stat, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice));
    h->roll = *cfg;
    h->have_roll = true;
    return GLG_OK;
}
extern "C" int glg_rollout_store(glg_handle *h, int32_t t, void *stream) {
    if (!h) return GLG_ERR_ARG;
    if (!h->have_roll) return fail(h, GLG_ERR_STATE, "glg_rollout_store: no rollout buffer (glg_rollout_create)");
    if (t < -1 || t >= h->roll.n_steps) return fail(h, GLG_ERR_ARG, "glg_rollout_store: t outside [-1, n_steps)");
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t B = (size_t)h->B, D = (size_t)h->obs_dim;
    GlgRollArgs a;
    memset(&a, 0, sizeof a);
    a.B = h->B; a.obs_dim = h->obs_dim; a.n_blocks = h->roll_blocks;
    a.training = h->roll.training; a.norm_obs = h->roll.norm_obs; a.norm_reward = h->roll.norm_reward;
    a.gamma = h->roll.gamma; a.clip_obs = h->roll.clip_obs; a.clip_reward = h->roll.clip_reward; a.epsilon = h->roll.epsilon;
    a.obs = h->obs;
    a.reward = t >= 0 ? h->reward : nullptr;
    a.done = t >= 0 ? h->done : nullptr;
    a.ret = h->roll_acc; a.stat = h->roll_stat; a.partial = h->roll_partial; a.norm64 = h->roll_norm;
    a.obs_out = h->roll_obs + (size_t)(t + 1) * B * D;
    a.reward_out = t >= 0 ? h->roll_rew + (size_t)t * B : nullptr;
    a.starts_out = h->roll_starts + (size_t)(t + 1) * B;
    cudaStream_t s = (cudaStream_t)stream;
    glg_roll_moments_kernel<<<h->roll_blocks, GLG_ROLL_NT, 0, s>>>(a);
    glg_roll_finish_kernel<<<(h->obs_dim + 1 + GLG_ROLL_NT / 32 - 1) / (GLG_ROLL_NT / 32), GLG_ROLL_NT, 0, s>>>(a);
    glg_roll_apply_kernel<<<h->roll_blocks, GLG_ROLL_NT, 0, s>>>(a);
    h->launches += 3;
    GLG_CUDA(h, cudaGetLastError());
    return GLG_OK;
}
extern "C" int glg_rollout_carry(glg_handle *h, void *stream) {
    if (!h) return GLG_ERR_ARG;
    if (!h->have_roll) return fail(h, GLG_ERR_STATE, "glg_rollout_carry: no rollout buffer (glg_rollout_create)");
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t B = (size_t)h->B, D = (size_t)h->obs_dim, T = (size_t)h->roll.n_steps;
    GLG_CUDA(h, cudaMemcpyAsync(h->roll_obs, h->roll_obs + T * B * D, B * D * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    GLG_CUDA(h, cudaMemcpyAsync(h->roll_starts, h->roll_starts + T * B, B * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return GLG_OK;
}
extern "C" int glg_rollout_gae(glg_handle *h, const float *values_dev, void *stream) {
    if (!h || !values_dev) return GLG_ERR_ARG;
    if (!h->have_roll) return fail(h, GLG_ERR_STATE, "glg_rollout_gae: no rollout buffer (glg_rollout_create)");
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    glg_roll_gae_kernel<<<(h->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->roll.n_steps, h->B, h->roll.gamma, h->roll.gae_lambda, h->roll_rew,
                                                                              values_dev, h->roll_starts, h->roll_adv, h->roll_ret);
    h->launches += 1;
    GLG_CUDA(h, cudaGetLastError());
    return GLG_OK;
}
extern "C" float *glg_rollout_obs_dev(glg_handle *h) { return h ? h->roll_obs : nullptr; }
extern "C" float *glg_rollout_rewards_dev(glg_handle *h) { return h ? h->roll_rew : nullptr; }
extern "C" float *glg_rollout_starts_dev(glg_handle *h) { return h ? h->roll_starts : nullptr; }
extern "C" float *glg_rollout_advantages_dev(glg_handle *h) { return h ? h->roll_adv : nullptr; }
extern "C" float *glg_rollout_returns_dev(glg_handle *h) { return h ? h->roll_ret : nullptr; }
extern "C" double *glg_rollout_stats_dev(glg_handle *h) { return h ? h->roll_stat : nullptr; }

// ---- episode-statistics all-reduce over NCCL (run-time binding: nccl_api above)
extern "C" int glg_nccl_unique_id(uint8_t id_out[128]) {
    if (!id_out) return GLG_ERR_ARG;
    std::string err;
    NcclApi *n = nccl_api(&err);
    if (!n) {
        g_create_error = err;
        return GLG_ERR_STATE;
    }
    NcclId id;
    const int rc = n->GetUniqueId(&id);
    if (rc != 0) {
        g_create_error = std::string("ncclGetUniqueId: ") + (n->GetErrorString ? n->GetErrorString(rc) : "error");
        return GLG_ERR_CUDA;
    }
    memcpy(id_out, id.internal, 128);
    return GLG_OK;
}
extern "C" int glg_nccl_init(glg_handle *h, const uint8_t id[128], int32_t rank, int32_t world_size) {
    if (!h || !id || world_size < 1 || rank < 0 || rank >= world_size) return GLG_ERR_ARG;
    NcclApi *n = nccl_api(&h->err);
    if (!n) return GLG_ERR_STATE;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    if (h->nccl_comm) {
        n->CommDestroy(h->nccl_comm);
        h->nccl_comm = nullptr;
    }
    NcclId nid;
    memcpy(nid.internal, id, 128);
    const int rc = n->CommInitRank(&h->nccl_comm, world_size, nid, rank);
    if (rc != 0) {
        h->nccl_comm = nullptr;
        return fail(h, GLG_ERR_CUDA, (std::string("ncclCommInitRank: ") + (n->GetErrorString ? n->GetErrorString(rc) : "error")).c_str());
    }
    return GLG_OK;
}
extern "C" int glg_allreduce_stats(glg_handle *h, void *stream) {
    if (!h) return GLG_ERR_ARG;
    if (!h->nccl_comm) return fail(h, GLG_ERR_STATE, "glg_allreduce_stats: no communicator (glg_nccl_init)");
    NcclApi *n = nccl_api(&h->err);
    if (!n) return GLG_ERR_STATE;
    GLG_CUDA(h, cudaSetDevice(h->cfg.device));
    const int rc = n->AllReduce(h->stats, h->stats, GLG_NSTATS, /*ncclDouble*/ 8, /*ncclSum*/ 0, h->nccl_comm, (cudaStream_t)stream);
    if (rc != 0) return fail(h, GLG_ERR_CUDA, (std::string("ncclAllReduce: ") + (n->GetErrorString ? n->GetErrorString(rc) : "error")).c_str());
    return GLG_OK;
}

// ---- batched evalF (handle-free) ------------------------------------------------------------------------
static thread_local std::string g_evalf_error;

template <bool GENERAL, bool PER_ENV_P>
static cudaError_t launch_evalf(const GlgUniform &uni, const double *x, const double *u, const double *d, const double *p,
                                double *xn, unsigned char *bad, int B, double dt, int n_sub, int integrator, cudaStream_t s) {
    constexpr int NT = 64;
    const size_t smem = sizeof(double) * (size_t)(2 * GLG_NX + H_COUNT) * NT;
    cudaError_t e = cudaFuncSetAttribute(glg_evalf_kernel<GENERAL, PER_ENV_P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    glg_evalf_kernel<GENERAL, PER_ENV_P, NT><<<(B + NT - 1) / NT, NT, smem, s>>>(uni, x, u, d, p, xn, bad, B, dt, n_sub, integrator);
    return cudaGetLastError();
}

extern "C" int glg_evalf_batch(const double *x_dev, const double *u_dev, const double *d_dev, const double *p_dev,
                               int32_t p_stride, double *x_next_dev, uint8_t *bad_dev, int32_t B, double dt, int32_t n_sub,
                               int32_t device, void *stream) {
    return glg_evalf_batch_ex(x_dev, u_dev, d_dev, p_dev, p_stride, x_next_dev, bad_dev, B, dt, n_sub, 0, device, stream);
}

extern "C" int glg_evalf_batch_ex(const double *x_dev, const double *u_dev, const double *d_dev, const double *p_dev,
                                  int32_t p_stride, double *x_next_dev, uint8_t *bad_dev, int32_t B, double dt, int32_t n_sub,
                                  int32_t integrator, int32_t device, void *stream) {
    if (!x_dev || !u_dev || !d_dev || !p_dev || !x_next_dev || B < 1 || n_sub < 1 || (p_stride != 0 && p_stride != GLG_NP) ||
        (integrator != 0 && integrator != 1)) {
        g_create_error = "glg_evalf_batch: invalid argument";
        return GLG_ERR_ARG;
    }
    cudaError_t e = cudaSetDevice(device);
    cudaStream_t s = (cudaStream_t)stream;
    GlgUniform uni{};  // only read by the shared-p variants (passed to the kernel by value)
    if (e == cudaSuccess) {
        if (p_stride == 0) {
            double ph[GLG_NP];
            e = cudaMemcpyAsync(ph, p_dev, sizeof ph, cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e == cudaSuccess) {
                for (int i = 0; i < GLG_NP; ++i) uni.P[i] = ph[i];
                glg_make_k(ph, uni.K);
                glg_make_c(ph, uni.C);
                e = glg_params_nominal_structure(ph)
                        ? launch_evalf<false, false>(uni, x_dev, u_dev, d_dev, p_dev, x_next_dev, bad_dev, B, dt, n_sub, integrator, s)
                        : launch_evalf<true, false>(uni, x_dev, u_dev, d_dev, p_dev, x_next_dev, bad_dev, B, dt, n_sub, integrator, s);
            }
        } else {
            e = launch_evalf<true, true>(uni, x_dev, u_dev, d_dev, p_dev, x_next_dev, bad_dev, B, dt, n_sub, integrator, s);
        }
    }
    if (e != cudaSuccess) {
        g_create_error = std::string("glg_evalf_batch: ") + cudaGetErrorString(e);
        return GLG_ERR_CUDA;
    }
    return GLG_OK;
}

// ---- roofline denominators ------------------------------------------------------------------------------
template <typename T>
static int measure_peak(int device, double *out) {
    if (!out) return GLG_ERR_ARG;
    cudaError_t e = cudaSetDevice(device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        g_create_error = std::string("glg_measure_peak: ") + cudaGetErrorString(e);
        return GLG_ERR_CUDA;
    }
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    T *buf = nullptr;
    cudaMalloc((void **)&buf, (size_t)blocks * threads * sizeof(T));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        glg_fma_peak_kernel<T><<<blocks, threads>>>(buf, iters, (T)0.999999, (T)1e-7);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads / (ms * 1e-3);
        if (rep > 0 && flops > best) best = flops;
    }
    e = cudaGetLastError();
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(buf);
    if (e != cudaSuccess) {
        g_create_error = std::string("glg_measure_peak: ") + cudaGetErrorString(e);
        return GLG_ERR_CUDA;
    }
    *out = best;
    return GLG_OK;
}
extern "C" int glg_measure_fp64_peak(int32_t device, double *flops) { return measure_peak<double>(device, flops); }
extern "C" int glg_measure_fp32_peak(int32_t device, double *flops) { return measure_peak<float>(device, flops); }

#ifdef GLG_TRACE
extern "C" int glg_debug_trace(long long *out_host) {  // [32 evaluations][16 warps][wake, arrive] clock64 stamps of CTA 0
    return cudaMemcpyFromSymbol(out_host, glg_trace, sizeof(long long) * 32 * 16 * 2) == cudaSuccess ? GLG_OK : GLG_ERR_CUDA;
}
#endif

extern "C" int glg_debug_math(int32_t op, const double *in_dev, double *out_dev, int32_t n, void *stream) {
    if (!in_dev || !out_dev || n < 1) return GLG_ERR_ARG;
    glg_math_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(op, in_dev, out_dev, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_create_error = std::string("glg_debug_math: ") + cudaGetErrorString(e);
        return GLG_ERR_CUDA;
    }
    return GLG_OK;
}
