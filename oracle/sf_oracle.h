/*
 * sf_oracle.h -- CPU restatement of the reference tick engine (TEST INFRASTRUCTURE ONLY).
 *
 * Plain-C model of the hot path of bistoyek21-ric/StrikeForce: random.hpp, the tick
 * functions of gameplay.hpp, the entity rules of Character.hpp / Item.hpp and the
 * observation builder of bots/bot-0.5/Custom.hpp.  Every function in sf_oracle.c cites the
 * reference file:line it follows.  It is pinned against the UNMODIFIED reference compiled
 * into oracle/_ref/libsfref.so (oracle/ref_harness) by tests/test_oracle_golden.py
 * and against the golden vectors under tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this.  The product (libstrikeforce_b200.so) never links or calls it.
 */
#ifndef SF_ORACLE_H
#define SF_ORACLE_H

#include <stdint.h>

#include "strikeforce_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sfo_arena sfo_arena;

/* one arena built from the same sf_config the product takes (n_envs, env_id_base ignored) */
sfo_arena *sfo_create(const sf_config *cfg);
void sfo_destroy(sfo_arena *a);

/* setup() + load_data() + _srand(tb, serial); gameplay.hpp:1231-1277, 1741-1925 */
void sfo_reset(sfo_arena *a, int level, int64_t tb, int64_t serial);
/* one env-step, gameplay.hpp:1443-1472; actions[i] = command symbol of human slot i */
int sfo_step(sfo_arena *a, const uint8_t *actions, int n);
/* the two halves of a step: A = spawns .. first update_bull, B = human_action .. end */
int sfo_step_a(sfo_arena *a);
int sfo_step_b(sfo_arena *a, const uint8_t *actions, int n);
int sfo_status(const sfo_arena *a);

/* frame kills teams_kills loot chest steps status hp */
void sfo_counters(const sfo_arena *a, int64_t out[8]);
/* humans zombies bullets chests built portals */
void sfo_population(const sfo_arena *a, int32_t out[6]);
long sfo_dump(const sfo_arena *a, int32_t *buf, long cap);
uint64_t sfo_hash(const sfo_arena *a);
/* sf_step_out-style deltas of the last step */
void sfo_step_out(const sfo_arena *a, sf_step_out *out);
/* draws consumed since reset (jomle delta) */
int64_t sfo_rng_draws(const sfo_arena *a);

/* gameplay::bot() up to predict(), bots/bot-0.5/Custom.hpp:137-158; returns SF_OBS_LEN or -1 */
int sfo_observe(const sfo_arena *a, int slot, float *out);
/* raw describe() planes without the pow transform, [32][31][31] */
int sfo_observe_raw(const sfo_arena *a, int slot, float *out);

/* random.hpp:54-76 on a stand-alone generator */
typedef struct sfo_rng { int64_t random[18], seed[18], us[18], jomle; } sfo_rng;
void sfo_srand(sfo_rng *r, int64_t tb, int64_t serial);
int sfo_rand(sfo_rng *r);
/* Character.hpp:29-45 */
int sfo_compute_damage(int x, int y);
/* the observation transform float(pow(double(fabsf(x) / 10), 0.2)), Custom.hpp:157 */
float sfo_obs_transform(float x);

/* free-running synthetic workload (sf_synth.h) with auto-reset; returns env-steps executed */
long sfo_run_stream(sfo_arena *a, int64_t env, int level, const char *table, int table_len,
                    long n_steps, int with_obs, uint64_t *hash_out);

/* the same workload with a checksum over the canonical-state hash after EVERY step
   (chk = mix64(chk ^ hash)), the last hash and optionally slot 0's observation at the end;
   returns the number of episodes that ended */
long sfo_run_trace(sfo_arena *a, int64_t env, int level, const char *table, int table_len, long n_steps,
                   uint64_t *chk_out, uint64_t *last_hash_out, float *obs_out);

#ifdef __cplusplus
}
#endif
#endif
