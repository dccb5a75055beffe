// Host build of the device tick (strikeforce_b200/csrc/sf_core.cuh) for CPU-side debugging.
//
// TEST INFRASTRUCTURE ONLY.  The product (libstrikeforce_b200.so) is CUDA-only and has no CPU
// path; this file compiles the very same per-arena functions with g++ so that
// `pytest -m "not gpu"` can diff the device algorithm against the CPU models in a container
// without a GPU.  Nothing in strikeforce_b200/ links or loads it.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sf_canon_dev.cuh"
#include "sf_core.cuh"
#include "sf_host_setup.h"
#include "sf_obs.cuh"

struct HcHandle {
    SfDev d;
    SfConst k;
    sfhost::Tables tabs;
    SfTabs t;
    std::vector<void *> allocs;
    unsigned long long stats[SF_STAT_COUNT];
    uint16_t bt[SF_BT_ENTRIES]; /* the bullet-flag table of whichever arena is being stepped */
    template <class T> T *alloc(size_t n)
    {
        void *p = calloc(n, sizeof(T));
        allocs.push_back(p);
        return (T *)p;
    }
};

extern "C" {

HcHandle *hc_create(const sf_config *cfg)
{
    HcHandle *h = new HcHandle();
    std::string err = sfhost::build_const(*cfg, h->k, h->tabs);
    if (!err.empty()) {
        fprintf(stderr, "hc_create: %s\n", err.c_str());
        delete h;
        return nullptr;
    }
    SfDev &d = h->d;
    const SfConst &k = h->k;
    size_t E = (size_t)((cfg->n_envs + 31) / 32 * 32);
    d.n_envs = cfg->n_envs, d.E = (int32_t)E;
    d.cap_t = (k.cap_t + 7) / 8 * 8;
    d.frame = h->alloc<uint32_t>(E), d.kills = h->alloc<int32_t>(E), d.tkills = h->alloc<int32_t>(E);
    d.loot = h->alloc<int32_t>(E), d.chest = h->alloc<int32_t>(E), d.misc = h->alloc<uint32_t>(E);
    d.steps = h->alloc<uint32_t>(E), d.episode = h->alloc<uint32_t>(E), d.ntemp = h->alloc<uint32_t>(E);
    d.mh = h->alloc<uint64_t>(E), d.mz = h->alloc<uint64_t>(2 * E), d.mb = h->alloc<uint64_t>(2 * E);
    d.mp = h->alloc<uint64_t>(2 * E);
    d.rng_log = h->alloc<uint32_t>(9 * E), d.rng_cst = h->alloc<uint32_t>(36 * E), d.jomle = h->alloc<uint32_t>(E);
    d.rng_w = h->alloc<uint32_t>(E);
    d.pend_log = h->alloc<uint32_t>(9 * E), d.pend_n = h->alloc<uint32_t>(E);
    size_t H = (size_t)k.cap_h * E, Z = (size_t)k.cap_z * E, B = (size_t)k.cap_b * E, T = (size_t)d.cap_t * E;
    d.h_pw = h->alloc<uint16_t>(H), d.h_sel = h->alloc<uint16_t>(H), d.h_bp = h->alloc<uint32_t>(H);
    d.h_hp = h->alloc<int32_t>(H), d.h_mind = h->alloc<int32_t>(H), d.h_stam = h->alloc<int32_t>(H);
    d.h_kills = h->alloc<int32_t>(H), d.h_dmg = h->alloc<int32_t>(H), d.h_eff = h->alloc<int32_t>(H);
    d.h_cons = h->alloc<uint32_t>(H), d.h_thr = h->alloc<uint32_t>(H);
    d.h_cmd = h->alloc<uint8_t>(H);
    d.z_pos = h->alloc<uint16_t>(Z), d.z_hp = h->alloc<int32_t>(Z), d.z_mind = h->alloc<int32_t>(Z);
    d.b_pw = h->alloc<uint16_t>(B), d.b_meta = h->alloc<uint32_t>(B), d.b_dmg = h->alloc<int32_t>(B);
    d.b_eff = h->alloc<int32_t>(B);
    d.t_cell = h->alloc<uint16_t>(T), d.t_dmg = h->alloc<int32_t>(T), d.t_pidx = h->alloc<uint8_t>(T);
    d.p_cell = h->alloc<uint16_t>((size_t)k.cap_p * E);
    d.grid = h->alloc<uint16_t>(E * SF_GRID_STRIDE);
    d.out = h->alloc<sf_step_out>(E);
    d.stats = h->stats;
    for (auto &s : h->stats) s = 0;
    d.smap = h->tabs.smap.data(), d.exp_tab = h->tabs.exp_tab.data(), d.log_tab = h->tabs.log_tab.data();
    sfhost::build_pow_lut(h->tabs, 1 << 16); /* small on purpose: exercises the beyond-table path */
    d.pow_lut = h->tabs.pow_lut.data(), d.pow_lut_len = (int32_t)h->tabs.pow_lut.size();
    h->t.smap = d.smap, h->t.exp_tab = d.exp_tab, h->t.log_tab = d.log_tab;
    h->t.rng_cst = d.rng_cst, h->t.E = d.E;
    h->t.bt = h->bt, h->t.bt_stride = 1;
    for (int env = 0; env < d.n_envs; ++env) {
        int64_t ge = k.env_id_base + env;
        sf_reset_body(d, k, h->t, env, sf_synth_tb(ge), sf_synth_serial(ge, 0), 0);
    }
    return h;
}

void hc_destroy(HcHandle *h)
{
    for (void *p : h->allocs) free(p);
    delete h;
}

void hc_reset(HcHandle *h, int env, long long tb, long long serial)
{
    sf_reset_body(h->d, h->k, h->t, env, tb, serial, 0);
}

static void add_stats(HcHandle *h, const SfStatDelta &sd)
{
    unsigned long long *s = h->stats;
    s[SF_STAT_STEPS] += sd.steps, s[SF_STAT_EPISODES] += sd.episodes, s[SF_STAT_WINS] += sd.wins;
    s[SF_STAT_DEATHS] += sd.deaths, s[SF_STAT_TIMEOUTS] += sd.timeouts, s[SF_STAT_TRUNCATED] += sd.truncated;
    s[SF_STAT_OVERFLOWS] += sd.overflows, s[SF_STAT_UB_GUARDS] += sd.ub_guards;
    s[SF_STAT_KILLS] += (long long)sd.kills, s[SF_STAT_TEAMS_KILLS] += (long long)sd.tkills;
    s[SF_STAT_LOOT] += (long long)sd.loot, s[SF_STAT_RNG_DRAWS] += sd.draws, s[SF_STAT_ALGO_BYTES] += sd.algo_bytes;
}

// actions: [n_envs][n_agents]
void hc_step(HcHandle *h, const uint8_t *actions, int half)
{
    SfStatDelta sd = {};
    for (int env = 0; env < h->d.n_envs; ++env)
        sf_step_body(h->d, h->k, h->t, env, true, actions ? actions + (size_t)env * h->k.n_agents : nullptr, half, sd);
    add_stats(h, sd);
}

long hc_dump(HcHandle *h, int env, int32_t *buf, long cap)
{
    SfBufSink sink{buf, cap, 0, false};
    sf_canon_emit(h->d, h->k, h->t, env, sink);
    return sink.overflow ? -1 : sink.n;
}

unsigned long long hc_hash(HcHandle *h, int env)
{
    SfHashSink sink{0};
    sf_canon_emit(h->d, h->k, h->t, env, sink);
    return sink.sum;
}

void hc_step_out(HcHandle *h, int env, sf_step_out *out) { *out = h->d.out[env]; }
int hc_status(HcHandle *h, int env) { return (int)((h->d.misc[env] >> 8) & 0xFF); }
unsigned hc_misc(HcHandle *h, int env) { return h->d.misc[env]; }
void hc_stats(HcHandle *h, unsigned long long *out)
{
    for (int i = 0; i < SF_STAT_COUNT; ++i) out[i] = h->stats[i];
}
// observation of human `slot` as bot() builds it (transformed), or the raw describe() planes
int hc_observe(HcHandle *h, int env, int slot, float *out, int raw)
{
    const SfDev &d = h->d;
    SfEnv e;
    sf_load_env(d, env, e);
    if (slot < 0 || slot >= e.hw_h) return -1;
    int vcell = (int)(SF_AT(d.h_pw, slot) & POS_CELL);
    uint32_t team = SF_AT(d.h_sel, slot) & HS_TEAM;
    uint32_t fb = 0;
    int32_t f[32];
    for (int wi = 0; wi < SF_OBS_WIN; ++wi)
        for (int wj = 0; wj < SF_OBS_WIN; ++wj) {
            sf_describe_milli(d, h->k, h->t, env, e, sf_obs_cell(vcell, wi, wj), team, -2, -2, f);
            for (int c = 0; c < 32; ++c)
                out[c * SF_OBS_CELLS + wi * SF_OBS_WIN + wj] =
                    raw ? (float)(f[c] / 1000.0) : sf_obs_transform(d, f[c], &fb);
        }
    return SF_OBS_LEN;
}

#ifdef SF_DEBUG_HIST
void hc_dbg_hist(unsigned long long *out)
{
    for (int i = 0; i < 256; ++i) out[i] = sf_dbg_hist[i];
}
#endif
int hc_n_agents(HcHandle *h) { return h->k.n_agents; }
int hc_compute_damage(int x, int y) { return sfhost::compute_damage(x, y); }
float hc_obs_transform_milli(int n) { return sfhost::obs_transform_milli(n); }

} // extern "C"
