// Headless driver around the UNMODIFIED reference tick engine -- test infrastructure only.
//
// This translation unit includes the reference's own headers (through the symlink farm
// that build_ref.sh creates under oracle/_ref/build/, nothing is copied) and calls the
// member functions of `struct gameplay` in the order of gameplay::play()
// (reference gameplay.hpp:1443-1472).  The functions of play() that block on a TTY
// (check_end :1102, get_my_action :939, render_it :1420) are not called; check_end's
// predicates are restated in eval_end() below with the frame clock of SURVEY.md 7.4#5.
//
// Exposed as a C ABI so that tests/ and bench.py can drive it through ctypes.
// One process == one arena (the reference keeps its state in globals, gameplay.hpp:37-55).

#include "selected_custom.hpp"   // -> reference bots/bot-0.5/Custom.hpp -> gameplay.hpp -> ...

#include "sf_canon.h"
#include "sf_synth.h"

#include <unistd.h>

#include <sstream>

using namespace Environment::Field;
namespace ECh = Environment::Character;
namespace EIt = Environment::Item;
namespace ERn = Environment::Random;

namespace {

struct Caps { int h, z, b, chest, built, portal; };

Caps  g_caps = {9000, 9000, 9000, 9000, 9000, 9000};
int   g_status = SF_RUNNING;
int   g_mode = SF_MODE_SOLO;
int   g_hw_h = 0;           // human slots ever used this episode
long  g_steps = 0;          // env-steps this episode
long  g_max_steps = 0;
bool  g_squad_agents = false;
std::string g_template;
bool  g_inited = false;

const char *mode_name(int m)
{
    return m == SF_MODE_SOLO ? "Solo" : m == SF_MODE_TIMER ? "Timer" : m == SF_MODE_ROYALE ? "Battle Royal" : "Squad";
}

void track_and_check_caps()
{
    int hi_h = -1, hi_z = -1, hi_b = -1, hi_p = -1;
    for (int i = 0; i < 4096 && i < H; ++i) {
        if (mh[i]) hi_h = i;
        if (mz[i]) hi_z = i;
        if (mb[i]) hi_b = i;
        if (active[i]) hi_p = i;
    }
    if (hi_h + 1 > g_hw_h) g_hw_h = hi_h + 1;
    if (hi_h >= g_caps.h || hi_z >= g_caps.z || hi_b >= g_caps.b || hi_p >= g_caps.portal ||
        g.chest > g_caps.chest || (long)g.temp.size() > g_caps.built)
        g_status = SF_OVERFLOW;
}

// update_bull (gameplay.hpp:1059-1100) snapshots the cell in front of EVERY live bullet
// without a bounds test; a bullet on row 0 heading up would index row -1.
bool update_bull_would_go_out_of_bounds()
{
    for (int i = 0; i < B && i < 4096; ++i)
        if (mb[i]) {
            std::vector<int> v = bull[i].get_cor();
            int d = bull[i].get_way() - 1;
            int r = v[1] + g.wdx[d], c = v[2] + g.wdy[d];
            if (r < 0 || r >= N || c < 0 || c >= M)
                return true;
        }
    return false;
}

bool rivals_dead()
{
    return g.rivals_are_dead();
}

// check_end (gameplay.hpp:1102-1229), offline branches only, wall clock replaced by
// the frame clock: 40 ms per frame (gameplay.hpp:1942) => level*300 s == level*7500 frames.
void eval_end()
{
    if (g_status != SF_RUNNING)
        return;
    if (g.online && rivals_dead()) { g_status = SF_WIN; return; }   // first test of check_end(), :1103
    if (hum[ind].get_Hp() <= 0) { g_status = SF_DEAD; return; }
    if (g_mode == SF_MODE_ROYALE) {
        // an online match ends in no other way
    } else if (g_mode == SF_MODE_TIMER) {
        if (g.frame >= g.level * 7500)
            g_status = (g.kills < g.level * 5) ? SF_TIMEOUT : SF_WIN;
    } else if (g_mode == SF_MODE_SOLO) {
        if (g.level * 5 <= g.kills) g_status = SF_WIN;
    } else {
        if (g.level * 10 <= g.teams_kills && rivals_dead()) g_status = SF_WIN;
    }
    if (g_status == SF_RUNNING && g_max_steps > 0 && g_steps >= g_max_steps)
        g_status = SF_TRUNCATED;
}

int action_index(char c)
{
    int act = 0;
    for (int i = 0; i < (int)g.action.size(); ++i)
        if (g.action[i] == c) act = i;
    return act;
}

} // namespace

extern "C" {

int sfref_init(const char *rundir)
{
    if (chdir(rundir) != 0)
        return -1;
    if (!g_inited) {
        EIt::download_items();
        ERn::make_p();
        g_inited = true;
    }
    return 0;
}

// caps: {humans, zombies, bullets, chests, built cells, portal slots}; NULL = reference caps
int sfref_reset(int mode, int level, long long tb, long long serial, const char *player_template,
                int squad_agents, const int *caps, long max_steps)
{
    if (!g_inited) return -1;
    if (caps) g_caps = Caps{caps[0], caps[1], caps[2], caps[3], caps[4], caps[5]};
    else g_caps = Caps{9000, 9000, 9000, 9000, 9000, 9000};
    g_mode = mode;
    g_template = player_template;
    g_squad_agents = squad_agents != 0;
    g_max_steps = max_steps;
    ECh::me = ECh::Human();
    ECh::me.build(false, "", g_template);
    g.manual = false;                 // => using_an_agent: load_data() calls prepare(me)
    g.replay_mode = false;
    g.enable_logging = false;
    g.mode = mode_name(mode);
    g.level = level;
    g.chest = 0;                      // fresh process semantics (setup() never clears it)
    g.setup();                        // parses ./map, places the players, seeds from time()
    ERn::_srand(tb, serial);          // re-seed: offline load_data draws nothing after seeding
    g.tb = tb;
    g.serial_number = serial;
    hum[ind].agent->slot = ind;
    if (g_squad_agents && mode == SF_MODE_SQUAD)
        for (int i = 1; i < 10; ++i) {          // USE_AGENT_IN_SQUAD_NPCS, gameplay.hpp:1886-1901
            g.prepare(hum[i]);
            hum[i].agent->slot = i;
        }
    ++g.frame;                        // play(): "++frame, find_recom(), render_it()" before the loop
    g_status = SF_RUNNING;
    g_steps = 0;
    g_hw_h = 0;
    std::memset(g_hub.captured, 0, sizeof(g_hub.captured));
    track_and_check_caps();
    return 0;
}

// Replay / logging variants of sfref_reset (reference gameplay.hpp:1749-1794, 966-993): the
// reference's OWN reader and writer of the .sf_sample format, used to pin strikeforce_b200/replay.py.
//   replay_path != NULL: load_data() reads seeds, player sheet and later every command from the
//                        file (the file name is asked on std::cin, so std::cin is redirected);
//   enable_logging:      load_data() opens ./datasets/.../(date).sf_sample and human_action()
//                        appends the commands; the seeds are the reference's own (time based),
//                        read them back with sfref_seeds().
int sfref_reset_ex(int mode, int level, const char *player_template, int squad_agents, const int *caps,
                   long max_steps, int enable_logging, const char *replay_path)
{
    if (!g_inited) return -1;
    if (caps) g_caps = Caps{caps[0], caps[1], caps[2], caps[3], caps[4], caps[5]};
    else g_caps = Caps{9000, 9000, 9000, 9000, 9000, 9000};
    g_mode = mode;
    g_template = player_template;
    g_squad_agents = squad_agents != 0;
    g_max_steps = max_steps;
    if (g.log_file.is_open()) g.log_file.close();
    if (g.replay_file.is_open()) g.replay_file.close();
    ECh::me = ECh::Human();
    ECh::me.build(false, "", g_template);
    g.manual = false;
    g.replay_mode = replay_path != nullptr;
    g.enable_logging = enable_logging != 0;
    g.log_filename = "";
    g.mode = mode_name(mode);
    g.level = level;
    g.chest = 0;
    std::streambuf *old_cin = std::cin.rdbuf();
    std::istringstream fake(std::string(replay_path ? replay_path : "") + "\n");
    std::cin.rdbuf(fake.rdbuf());
    std::streambuf *old_cout = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf()); // the prompt "Enter the file's address: "
    g.setup();
    std::cin.rdbuf(old_cin);
    std::cout.rdbuf(old_cout);
    hum[ind].agent->slot = ind;
    if (g_squad_agents && mode == SF_MODE_SQUAD)
        for (int i = 1; i < 10; ++i) {
            g.prepare(hum[i]);
            hum[i].agent->slot = i;
        }
    if (mode == SF_MODE_ROYALE)
        // Battle Royale through the reference's replay reader (gameplay.hpp:1762-1806, 1847-1859):
        // the other players are `remote`, their commands come from the file; giving them an oracle
        // agent changes nothing in the tick (human_action takes the remote branch, :981) and lets
        // sfref_observe show what each player's own client would see
        for (int i = 0; i < 64; ++i)
            if (i != ind && mh[i]) { // right after load_data() the live humans are the players
                g.prepare(hum[i]);
                hum[i].agent->slot = i;
            }
    ++g.frame;
    g_status = SF_RUNNING;
    g_steps = 0;
    g_hw_h = 0;
    std::memset(g_hub.captured, 0, sizeof(g_hub.captured));
    track_and_check_caps();
    return 0;
}

// A LIVE online match (gameplay.hpp:1807-1859): load_data() asks for the server's address, port and
// password on std::cin, and the reference's own client code (gameplay.hpp:66-193) connects, receives
// seeds and roster, announces the account sheet of `user` (./accounts/game/<user>/info, <user>.txt in
// the run directory) and receives the other players'.  Used to pin strikeforce_b200/match_server.py.
int sfref_reset_online(const char *ip, int port, const char *password, const char *user_name, const int *caps,
                       long max_steps)
{
    if (!g_inited) return -1;
    if (caps) g_caps = Caps{caps[0], caps[1], caps[2], caps[3], caps[4], caps[5]};
    else g_caps = Caps{9000, 9000, 9000, 9000, 9000, 9000};
    g_mode = SF_MODE_ROYALE;
    g_squad_agents = false;
    g_max_steps = max_steps;
    if (g.log_file.is_open()) g.log_file.close();
    if (g.replay_file.is_open()) g.replay_file.close();
    user = user_name;
    g_template = std::string("./accounts/game/") + user_name + "/info, " + user_name + ".txt";
    ECh::me = ECh::Human();
    ECh::me.build(false, "", g_template);
    g.manual = false;
    g.replay_mode = false;
    g.enable_logging = false;
    g.mode = mode_name(SF_MODE_ROYALE);
    g.level = 1;
    g.chest = 0;
    std::streambuf *old_cin = std::cin.rdbuf();
    std::istringstream fake(std::string(ip) + "\n" + std::to_string(port) + "\n" + password + "\n");
    std::cin.rdbuf(fake.rdbuf());
    std::streambuf *old_cout = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    g.setup();
    std::cin.rdbuf(old_cin);
    std::cout.rdbuf(old_cout);
    if (disconnect) return -2;
    hum[ind].agent->slot = ind;
    ++g.frame;
    g_status = SF_RUNNING;
    g_steps = 0;
    g_hw_h = 0;
    std::memset(g_hub.captured, 0, sizeof(g_hub.captured));
    track_and_check_caps();
    return ind;
}

// the seeds of the running match (after a logging reset: the reference's own, as written to the log)
void sfref_seeds(long long *out) { out[0] = (long long)g.tb; out[1] = (long long)g.serial_number; }

// close the log of a logging match and return its path (relative to the run directory)
int sfref_close_log(char *out, int cap)
{
    if (g.log_file.is_open()) g.log_file.close();
    g.enable_logging = false;
    std::snprintf(out, cap, "%s", g.log_filename.c_str());
    return (int)g.log_filename.size();
}

int sfref_status() { return g_status; }

// actions[i] = command symbol for human slot i (slot 0 = the player).  Agent-driven squad
// NPCs can only emit the 9 symbols of gameplay::action (bots/bot-0.5/Custom.hpp:162).
int sfref_step(const unsigned char *actions, int n)
{
    if (g_status != SF_RUNNING) return g_status;
#ifdef SFREF_NO_GUARDS /* timing build (libsfref_noguard.so): what the harness's own checks cost the CPU baseline */
#define SF_CHECK() ((void)0)
#define SF_UBGUARD() ((void)0)
#else
#define SF_CHECK() do { track_and_check_caps(); if (g_status != SF_RUNNING) return g_status; } while (0)
#define SF_UBGUARD() do { if (update_bull_would_go_out_of_bounds()) { g_status = SF_UB_GUARD; return g_status; } } while (0)
#endif
    if (g.frame % g.pc <= 1) g.spawn_chest();
    if (g.frame % g.pz <= 1) g.spawn_zombie_npc();
    if (g.frame % g.ph <= 1) g.spawn_human_npc();
    SF_CHECK();
    command[ind] = n > 0 ? (char)actions[0] : '+';
    if (g.online && !g.replay_mode) client.send_it();   // get_my_action(), gameplay.hpp:960-961
    for (int i = 1; i < 64 && i < n; ++i)
        g_hub.next_action[i] = action_index((char)actions[i]);
    g.zombie_action();
    SF_CHECK();
    g.portal_damage();
    SF_CHECK();
    g.view();
    g.update_tmp();
    g.hit_human(), g.hit_zombie();
    ++g.frame;
    g.updmap();
    SF_UBGUARD();
    g.update_bull();
    g.human_action();
    SF_CHECK();
    g.view();
    g.update_tmp();
    g.hit_human(), g.hit_zombie();
    ++g.frame;
    g.updmap();
    SF_UBGUARD();
    g.update_bull();
    ++g_steps;
    eval_end();
    if (g.online && !g.replay_mode && (g_status == SF_DEAD || g_status == SF_WIN)) {
        command[ind] = g_status == SF_DEAD ? '~' : '+';   // check_end(), gameplay.hpp:1103-1108, 1131-1136
        client.send_it();
        client.end_it();
    }
    return g_status;
#undef SF_CHECK
#undef SF_UBGUARD
}

// P1 observation of human `slot` at the current state (get_my_action, gameplay.hpp:956).
int sfref_observe(int slot, float *out)
{
    if (slot < 0 || slot >= 64 || !hum[slot].get_active_agent()) return -1;
    int save_cap = g_hub.capture[slot], save_cnt = g_hub.captured[slot];
    g_hub.capture[slot] = 1;
    g.bot(hum[slot]);
    g_hub.capture[slot] = save_cap;
    g_hub.captured[slot] = save_cnt;
    std::memcpy(out, g_hub.obs[slot], sizeof(float) * OracleAgentHub::OBS_LEN);
    return OracleAgentHub::OBS_LEN;
}

// P2 observations (inside human_action, gameplay.hpp:933): enable capture before a step,
// fetch afterwards.
void sfref_set_capture(int slot, int on) { if (slot >= 0 && slot < 64) g_hub.capture[slot] = on, g_hub.captured[slot] = 0; }
int sfref_get_captured(int slot, float *out)
{
    if (slot < 0 || slot >= 64 || !g_hub.captured[slot]) return 0;
    std::memcpy(out, g_hub.obs[slot], sizeof(float) * OracleAgentHub::OBS_LEN);
    return OracleAgentHub::OBS_LEN;
}

// counters: frame kills teams_kills loot chest steps status Hp(main)
void sfref_counters(long long *out)
{
    out[0] = g.frame; out[1] = g.kills; out[2] = g.teams_kills; out[3] = g.loot;
    out[4] = g.chest; out[5] = g_steps; out[6] = g_status; out[7] = hum[ind].get_Hp();
}

// live populations: humans zombies bullets chests built portals
void sfref_population(int *out)
{
    int nh = 0, nz = 0, nb = 0, np = 0;
    for (int i = 0; i < 4096; ++i) { nh += mh[i]; nz += mz[i]; nb += mb[i]; np += active[i]; }
    out[0] = nh; out[1] = nz; out[2] = nb; out[3] = (int)g.chest; out[4] = (int)g.temp.size(); out[5] = np;
}

static int emit(int32_t *buf, long cap, long &n, int kind, int index, const int32_t *f, int nf)
{
    if (n + 3 + nf > cap) return -1;
    buf[n++] = kind; buf[n++] = index; buf[n++] = nf;
    for (int i = 0; i < nf; ++i) buf[n++] = f[i];
    return 0;
}

// canonical record (include/sf_canon.h); returns the number of int32 written, <0 if cap too small
long sfref_dump(int32_t *buf, long cap)
{
    long n = 0;
    int32_t f[32];
    f[0] = g_mode; f[1] = (int)g.level; f[2] = (int)g.frame; f[3] = (int)g.kills; f[4] = (int)g.teams_kills;
    f[5] = (int)g.loot; f[6] = (int)g.chest; f[7] = ind;
    if (emit(buf, cap, n, SF_K_HEADER, 0, f, SF_NF_HEADER)) return -1;
    for (int i = 0; i < 18; ++i) f[i] = (int)ERn::random[i];
    f[18] = (int)(ERn::jomle & 0xFFFF);
    if (emit(buf, cap, n, SF_K_RNG, 0, f, SF_NF_RNG)) return -1;
    for (int i = 0; i < g_hw_h; ++i) {
        const ECh::Human &h = hum[i];
        std::vector<int> v = h.get_cor();
        int k = 0;
        f[k++] = mh[i]; f[k++] = h.is_rnpc(); f[k++] = h.get_team(); f[k++] = h.get_way();
        f[k++] = v[0]; f[k++] = v[1]; f[k++] = v[2];
        f[k++] = h.get_Hp(); f[k++] = h.get_mindamage(); f[k++] = h.get_stamina();
        f[k++] = h.get_kills(); f[k++] = h.get_damage(); f[k++] = h.get_effect();
        f[k++] = h.backpack.vec; f[k++] = h.backpack.ind;
        for (int j = 0; j < 4; ++j) f[k++] = h.backpack.list_cons[j].second;
        for (int j = 0; j < 4; ++j) f[k++] = h.backpack.list_throw[j].second.second;
        f[k++] = h.backpack.get_blocks(); f[k++] = h.backpack.get_portals(); f[k++] = h.backpack.get_portal_ind();
        f[k++] = h.get_mindamage_def();
        if (emit(buf, cap, n, SF_K_HUMAN, i, f, SF_NF_HUMAN)) return -1;
    }
    for (int i = 0; i < Z && i < 4096; ++i)
        if (mz[i]) {
            std::vector<int> v = zomb[i].get_cor();
            f[0] = zomb[i].is_super(); f[1] = v[0]; f[2] = v[1]; f[3] = v[2];
            f[4] = zomb[i].get_Hp(); f[5] = zomb[i].get_mindamage();
            if (emit(buf, cap, n, SF_K_ZOMBIE, i, f, SF_NF_ZOMBIE)) return -1;
        }
    for (int i = 0; i < B && i < 4096; ++i)
        if (mb[i]) {
            std::vector<int> v = bull[i].get_cor(), d = bull[i].get_dcor();
            ECh::Human *o = reinterpret_cast<ECh::Human *>(bull[i].get_owner());
            f[0] = v[0]; f[1] = v[1]; f[2] = v[2]; f[3] = d[0]; f[4] = d[1]; f[5] = d[2];
            f[6] = bull[i].get_way(); f[7] = bull[i].get_range(); f[8] = bull[i].get_damage();
            f[9] = bull[i].get_effect(); f[10] = o ? (int)(o - hum) : -1;
            if (emit(buf, cap, n, SF_K_BULLET, i, f, SF_NF_BULLET)) return -1;
        }
    for (int i = 0; i < B && i < 4096; ++i)
        if (active[i]) {
            f[0] = portal[i][0]; f[1] = portal[i][1]; f[2] = portal[i][2];
            if (emit(buf, cap, n, SF_K_PORTAL, i, f, SF_NF_PORTAL)) return -1;
        }
    for (int a = 0; a < F; ++a)
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < M; ++c) {
                const node &x = g.themap[a][r][c];
                if (!(x.s[0] || x.s[1] || x.s[2] || x.s[4] || x.s[10])) continue;
                int kind = x.s[10] ? (x.s[3] ? 1 : x.s[5] ? 2 : x.s[7] ? 3 : 0) : 0;
                f[0] = x.s[0]; f[1] = x.s[0] ? (int)(x.human - hum) : -1;
                f[2] = x.s[1]; f[3] = x.s[1] ? (int)(x.zombie - zomb) : -1;
                f[4] = x.s[2]; f[5] = x.s[2] ? (int)(x.bullet - bull) : -1;
                f[6] = x.s[4]; f[7] = x.s[4] ? (int)(x.cons - EIt::cons) : -1;
                f[8] = kind; f[9] = x.dmg; f[10] = kind == 2 ? x.portal_ind : -1;
                if (emit(buf, cap, n, SF_K_CELL, (a * N + r) * M + c, f, SF_NF_CELL)) return -1;
            }
    return n;
}

unsigned long long sfref_hash()
{
    static int32_t buf[1 << 20];
    long n = sfref_dump(buf, 1 << 20);
    return n < 0 ? 0ULL : sf_canon_hash(buf, n);
}

// raw RNG access for the known-answer tests (random.hpp:54-76)
void sfref_srand(long long tb, long long serial) { if (g_inited) ERn::_srand(tb, serial); }
int  sfref_rand() { return ERn::_rand(); }
void sfref_rng_state(long long *out) { for (int i = 0; i < 18; ++i) out[i] = ERn::random[i]; out[18] = ERn::jomle; }
int  sfref_compute_damage(int x, int y) { return ECh::compute_damage(x, y); }

// Free-running loop for the CPU baseline: arena `env` of the synthetic workload
// (sf_synth.h), auto-reset on terminal status.  Returns env-steps executed.
// with_obs: also build the player's observation each step through the reference's bot().
long sfref_run_stream(long long env, int mode, int level, const char *player_template, int squad_agents,
                      const int *caps, long max_steps, const char *table, int table_len,
                      long n_steps, int with_obs, unsigned long long *hash_out)
{
    static float obs[OracleAgentHub::OBS_LEN];
    long long episode = 0;
    uint64_t streams[10];
    for (int a = 0; a < 10; ++a) streams[a] = sf_synth_stream_init(env, a);
    int n_agents = (mode == SF_MODE_SQUAD && squad_agents) ? 10 : 1;
    if (sfref_reset(mode, level, sf_synth_tb(env), sf_synth_serial(env, episode), player_template,
                    squad_agents, caps, max_steps)) return -1;
    unsigned char act[10];
    unsigned long long acc = 0;
    for (long s = 0; s < n_steps; ++s) {
        for (int a = 0; a < 10; ++a) {
            uint64_t z = sf_synth_stream_next(&streams[a]);
            act[a] = (unsigned char)table[z % (uint64_t)table_len];
        }
        if (with_obs) { sfref_observe(0, obs); acc += (unsigned long long)(obs[15 * 31 + 15] * 1000.f); }
        int st = sfref_step(act, n_agents);
        if (st != SF_RUNNING) {
            ++episode;
            sfref_reset(mode, level, sf_synth_tb(env), sf_synth_serial(env, episode), player_template,
                        squad_agents, caps, max_steps);
        }
    }
    if (hash_out) *hash_out = sfref_hash() + acc;
    return n_steps;
}

} // extern "C"
