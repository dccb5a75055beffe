/*
 * sf_oracle.c -- CPU restatement of the reference tick engine.  TEST INFRASTRUCTURE ONLY
 * (see sf_oracle.h).  All citations are file:line under
 * /root/reference/StrikeForce-client/.  "harness" = oracle/ref_harness/harness.cpp, the
 * headless driver of the unmodified reference whose conventions (frame clock, capacity
 * overflow, out-of-bounds guard, end-of-step victory test) this file shares so that the two
 * can be compared record for record.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py compares the canonical record
 * (include/sf_canon.h) of this model with oracle/_ref/libsfref.so after every step over
 * seeded matches (test_oracle_matches_reference_live; tests/test_royale_reference_live.py for
 * Battle Royale), and tests/golden/ holds records produced by the reference itself
 * (tests/golden/make_golden.py, make_golden_royale.py).
 */
#include "sf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "sf_synth.h"

#define F SF_FLOORS
#define N SF_ROWS
#define M SF_COLS
#define MAXS 1024          /* slot-array size of the model; configured caps must be smaller */
#define REF_C 9000         /* gameplay.hpp:37  C  */
#define LIM_PORTAL 1000    /* gameplay.hpp:37 */
#define LIM_BLOCK 1100
#define PC 30              /* gameplay.hpp:459 spawn periods */
#define PZ 40
#define PH 50

static const int wdx[4] = {1, 0, -1, 0}, wdy[4] = {0, 1, 0, -1}; /* gameplay.hpp:459 */

/* gameplay.hpp:237-243 -- pointers become slot indices (-1 = nullptr) */
typedef struct ocell {
    uint8_t s[11];
    int dmg, portal_ind;
    int human, zombie, bullet, cons;
} ocell;

typedef struct owpn { int stamina, damage, effect, range; } owpn;

/* Character.hpp:225-331 + Backpack :49-223; only what the tick reads or writes */
typedef struct ohuman {
    int rnpc, agent_active;
    int way, team, kills, damage, effect;
    int Hp, mindamage, mindamage_def, def_Hp, stamina, def_stamina;
    int cor[3];
    int vec, ind;
    int cons[4];
    int throw_lvl[4], throw_cnt[4];
    int w_lvl[8];
    int blocks, portals, portal_ind, def_blocks, def_portals;
} ohuman;

typedef struct ozombie { int super_, Hp, mindamage, cor[3]; } ozombie;

/* Item.hpp:120-172 */
typedef struct obullet { int way, cor[3], dcor[3], owner, damage, effect, range, stamina; } obullet;

struct sfo_arena {
    /* configuration */
    int mode, squad_agents, max_steps;
    int n_players, teams[SF_MAX_PLAYERS]; /* SF_MODE_ROYALE: `players` and the teams of the replay header */
    int32_t royale_sheets[SF_MAX_PLAYERS][SF_SHEET_LEN]; /* the sheets the other players announced */
    int royale_ind;                                       /* which player is `ind` in a Battle Royale match */
    int cap_h, cap_z, cap_b, cap_chest, cap_built, cap_portal;
    uint8_t map_cells[SF_CELLS];
    int16_t map_portal[SF_CELLS];
    sf_consumable cons[4];
    sf_weapon thr[4], wpn[8];
    int32_t player_sheet[SF_SHEET_LEN], npc_sheet[SF_SHEET_LEN];
    /* random.hpp:29-31 */
    sfo_rng rng;
    int64_t jomle0;
    /* gameplay.hpp:47-55 */
    int ind;
    ohuman hum[MAXS];
    ozombie zomb[MAXS];
    obullet bull[MAXS];
    int portal[MAXS][3];
    uint8_t mb[MAXS], active[MAXS], mz[MAXS], mh[MAXS];
    uint8_t remote[MAXS]; /* gameplay.hpp:53: slots of the other players of an online match */
    uint8_t command[MAXS];
    /* gameplay.hpp:459-475 */
    int64_t loot, level, teams_kills, kills, chest, frame;
    ocell themap[F][N][M];
    uint8_t map1_s2[F][N][M];  /* themap1: only s[2] and the bullet pointer are ever read back */
    int map1_bullet[F][N][M];
    int temp[SF_CELLS], n_temp;
    /* harness bookkeeping */
    int status, hw_h;
    long steps;
    /* deltas of the last step */
    sf_step_out out;
};

/* ------------------------------------------------------------------ random.hpp */

/* random.hpp:42-52 */
static int64_t binpow_(int64_t a, int64_t b)
{
    const int64_t mod = 65537;
    int64_t res = 1;
    b %= mod - 1;
    while (b) {
        if (b & 1) res = (res * a) % mod;
        a = (a * a) % mod;
        b >>= 1;
    }
    return res;
}

/* random.hpp:54-62; p[x][e] (make_p :33-40) is x^e mod 65537, computed directly here */
int sfo_rand(sfo_rng *r)
{
    const int64_t mod = 65537;
    int64_t sum = 1;
    for (int i = 0; i < 18; ++i) {
        int64_t pw = 1;
        for (int64_t j = 0; j < r->seed[i]; ++j) pw = (r->random[i] * pw) % mod;
        sum = (sum + r->us[i] * pw) % mod;
    }
    r->random[0] = binpow_(sum + (sum == 0), ++r->jomle);
    for (int i = 0; i < 17; ++i) {
        int64_t t = r->random[i];
        r->random[i] = r->random[i + 1];
        r->random[i + 1] = t;
    }
    return (int)(r->random[17] & 1023);
}

/* random.hpp:64-76 */
void sfo_srand(sfo_rng *r, int64_t tb, int64_t u_s)
{
    for (int i = 0; i < 18; ++i) {
        r->us[i] = u_s % 10 + 1;
        r->seed[i] = tb % 10 + 1;
        u_s /= 10;
        tb /= 10;
        r->random[i] = 0;
    }
    r->jomle = 18;
    for (int i = 0; i < 1024; ++i) sfo_rand(r);
}

#define RAND(a) sfo_rand(&(a)->rng) /* gameplay.hpp:33-35: Field::rand() is _rand() */

/* ------------------------------------------------------------------ Character.hpp */

/* Character.hpp:29-45 (the indentation there misleads: only `tmp /= mid` is in the for) */
int sfo_compute_damage(int x, int y)
{
    int l = 0, r = x + 1, z = 2;
    while (1 < y) {
        y >>= 1;
        ++z;
    }
    while (r - l > 1) {
        int mid = (l + r) >> 1, tmp = x;
        for (int i = 0; i < z && mid; ++i) tmp /= mid;
        if (tmp) l = mid;
        else r = mid;
    }
    return l;
}

static int imax(int a, int b) { return a > b ? a : b; }

/* level_solo_up / level_timer_up / level_squad_up, Character.hpp:765-801 */
static void level_up(ohuman *h, int *lvl)
{
    ++*lvl;
    h->mindamage_def += 5;
    h->def_Hp += 50;
    h->def_stamina += 50;
    if (*lvl % 2 == 1) {
        ++h->def_blocks;
        ++h->def_portals;
    }
}

/* Human::build, Character.hpp:650-709, with Backpack::build :74-88 and back_tmp :157-162.
 * The sheet is the file content after the name. */
static void human_build(ohuman *h, const int32_t *sheet, int rnpc)
{
    int lv[3];
    h->def_blocks = 8;
    h->def_portals = 1;
    h->portal_ind = -1;
    h->vec = h->ind = -1;
    h->rnpc = rnpc;
    h->def_Hp = sheet[0];
    h->mindamage_def = sheet[1];
    h->def_stamina = sheet[2];
    lv[0] = sheet[3];
    lv[1] = sheet[4];
    lv[2] = sheet[5];
    h->Hp = h->def_Hp, h->mindamage = h->mindamage_def, h->stamina = h->def_stamina;
    for (int i = 0; i < 4; ++i) h->cons[i] = sheet[11 + i];
    for (int i = 0; i < 4; ++i) {
        h->throw_lvl[i] = sheet[15 + 2 * i];
        h->throw_cnt[i] = sheet[16 + 2 * i];
    }
    for (int i = 0; i < 8; ++i) h->w_lvl[i] = sheet[23 + i];
    for (int m = 0; m < 3; ++m) {
        int k = lv[m], cur = 1;
        while (--k) level_up(h, &cur);
    }
    h->blocks = h->def_blocks;
    h->portals = h->def_portals;
    h->portal_ind = -1;
}

/* gen_human, Character.hpp:873-888 */
static void gen_human(sfo_arena *a, int rnpc, ohuman *h, int lvl, int f, int r, int c)
{
    int cur[3] = {a->npc_sheet[3], a->npc_sheet[4], a->npc_sheet[5]};
    h->cor[0] = f, h->cor[1] = r, h->cor[2] = c;
    h->way = 1;
    human_build(h, a->npc_sheet, 1);
    h->rnpc = rnpc;
    h->team = 0;
    h->kills = 0;
    h->damage = 0;
    h->effect = 0;
    while (--lvl) {
        level_up(h, &cur[0]);
        level_up(h, &cur[1]);
        level_up(h, &cur[2]);
    }
}

/* item stats as carried in a backpack: Weapon::upgrade() per level, Item.hpp:105-111,
 * applied lvl times to weapons (Character.hpp:683-686) and lvl-1 times to throwables (:676-680) */
static owpn weapon_of(const sfo_arena *a, const ohuman *h, int i)
{
    owpn w = {a->wpn[i].stamina, a->wpn[i].damage + 50 * h->w_lvl[i], a->wpn[i].effect - 50 * h->w_lvl[i],
              a->wpn[i].range};
    return w;
}
static owpn throw_of(const sfo_arena *a, const ohuman *h, int i)
{
    int up = h->throw_lvl[i] - 1 > 0 ? h->throw_lvl[i] - 1 : 0;
    owpn w = {a->thr[i].stamina, a->thr[i].damage + 50 * up, a->thr[i].effect - 50 * up, a->thr[i].range};
    return w;
}

/* Bullet::shot, Item.hpp:153-160 */
static void bullet_shot(obullet *b, int f, int r, int c, int way, int damage, int effect, int range, int owner)
{
    b->owner = owner;
    b->way = way;
    b->cor[0] = b->dcor[0] = f;
    b->cor[1] = b->dcor[1] = r;
    b->cor[2] = b->dcor[2] = c;
    b->damage = damage, b->effect = effect, b->range = range;
}

/* Human::punch, Character.hpp:391-397 */
static int human_punch(ohuman *h, int hidx, obullet *b)
{
    int d = imax(sfo_compute_damage(h->mindamage_def, 1), h->mindamage);
    bullet_shot(b, h->cor[0], h->cor[1] + wdx[h->way - 1], h->cor[2] + wdy[h->way - 1], h->way, d, 0, 1, hidx);
    return 1;
}

/* Human::shot_it, Character.hpp:399-408 */
static int human_shot_it(const sfo_arena *a, ohuman *h, int hidx, obullet *b)
{
    owpn w = weapon_of(a, h, h->ind);
    if (h->stamina + w.stamina < 0) return 0;
    h->stamina += w.stamina;
    w.damage = imax(sfo_compute_damage(w.damage, w.range), w.damage + h->mindamage);
    bullet_shot(b, h->cor[0], h->cor[1] + wdx[h->way - 1], h->cor[2] + wdy[h->way - 1], h->way, w.damage, w.effect,
                w.range, hidx);
    return 1;
}

/* Human::throw_it, Character.hpp:410-427 */
static int human_throw_it(const sfo_arena *a, ohuman *h, int hidx, obullet *b)
{
    owpn t = throw_of(a, h, h->ind);
    t.damage = imax(t.damage, t.damage + h->mindamage);
    if (h->stamina + t.stamina < 0) return 0;
    if (h->throw_cnt[h->ind] < 1) {
        h->vec = -1;
        return 0;
    }
    h->stamina += t.stamina;
    --h->throw_cnt[h->ind];
    if (h->throw_cnt[h->ind] < 1) h->vec = -1;
    bullet_shot(b, h->cor[0], h->cor[1] + wdx[h->way - 1], h->cor[2] + wdy[h->way - 1], h->way, t.damage, t.effect,
                t.range, hidx);
    return 1;
}

/* Human::use, Character.hpp:379-389 */
static void human_use(const sfo_arena *a, ohuman *h)
{
    if (h->vec || h->cons[h->ind] < 1) return;
    h->stamina += a->cons[h->ind].stamina;
    h->Hp += a->cons[h->ind].hp;
    h->mindamage += a->cons[h->ind].effect;
    if ((--h->cons[h->ind]) < 1) h->vec = -1;
}

/* Human::get_damage_effect, Character.hpp:429-443 */
static void human_damage_effect(const sfo_arena *a, const ohuman *h, int out[2])
{
    int dmg = imax(sfo_compute_damage(h->mindamage_def, 1), h->mindamage);
    if (h->vec == 1) {
        owpn t = throw_of(a, h, h->ind);
        if (0 <= h->stamina + t.stamina) {
            out[0] = imax(imax(t.damage, t.damage + h->mindamage), dmg);
            out[1] = t.effect;
            return;
        }
    }
    if (h->vec == 2) {
        owpn w = weapon_of(a, h, h->ind);
        if (0 <= h->stamina + w.stamina) {
            out[0] = imax(imax(sfo_compute_damage(w.damage, w.range), w.damage + h->mindamage), dmg);
            out[1] = w.effect;
            return;
        }
    }
    out[0] = dmg;
    out[1] = 0;
}

/* ------------------------------------------------------------------ gameplay.hpp */

/* node::showit, gameplay.hpp:321-341.  Humans print a direction glyph and zombies z/Z; no
 * caller distinguishes them from each other, so 'H' and 'z' stand in. */
static char showit(const ocell *x)
{
    if (x->s[3]) return '#';
    if (x->s[0]) return 'H';
    if (x->s[1]) return 'z';
    if (x->s[5]) return '^';
    if (x->s[6]) return 'v';
    if (x->s[2]) return '*';
    if (x->s[4]) return '?';
    if (x->s[8]) return 'X';
    if (x->s[7]) return 'O';
    return '.';
}

/* gameplay.hpp:209-235 */
static int p_ind(const sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i)
        if (!a->active[i]) return i;
    return -1;
}
static int h_ind(const sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i)
        if (i != a->ind && !a->mh[i] && !a->remote[i]) return i;
    return -1;
}
static int z_ind(const sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i)
        if (!a->mz[i]) return i;
    return -1;
}
static int b_ind(const sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i)
        if (!a->mb[i]) return i;
    return -1;
}

/* gameplay.hpp:489-495 */
static void updmap(sfo_arena *a)
{
    for (int i = 0; i < F; ++i)
        for (int j = 0; j < N; ++j)
            for (int k = 0; k < M; ++k) a->themap[i][j][k].s[8] = a->themap[i][j][k].s[9] = 0;
}

/* gameplay.hpp:497-505 */
static int rivals_are_dead(const sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i)
        if (a->mh[i]) {
            int team = a->hum[i].team;
            if (team && team != a->hum[a->ind].team) return 0;
        }
    return 1;
}

/* gameplay.hpp:507-515 with Human::claim_chest, Character.hpp:372-377 */
static void claim_chest(sfo_arena *a, ohuman *p)
{
    ocell *x = &a->themap[p->cor[0]][p->cor[1]][p->cor[2]];
    if (x->s[4]) {
        p->stamina += a->cons[x->cons].stamina;
        p->Hp += a->cons[x->cons].hp;
        p->mindamage += a->cons[x->cons].effect;
        x->s[4] = 0;
        --a->chest;
    }
}

/* gameplay.hpp:517-530 */
static void teleport(sfo_arena *a, int hidx)
{
    ohuman *p = &a->hum[hidx];
    ocell *src = &a->themap[p->cor[0]][p->cor[1]][p->cor[2]];
    int index = src->portal_ind;
    if (index == -1) return;
    ocell *dst = &a->themap[a->portal[index][0]][a->portal[index][1]][a->portal[index][2]];
    if (showit(dst) != 'O') return;
    dst->s[0] = 1;
    dst->human = hidx;
    src->s[0] = 0;
    p->cor[0] = a->portal[index][0], p->cor[1] = a->portal[index][1], p->cor[2] = a->portal[index][2];
}

/* gameplay.hpp:532-542 */
static void spawn_chest(sfo_arena *a)
{
    if (REF_C <= a->chest) return;
    int i = RAND(a) % F, j = RAND(a) % N, k = RAND(a) % M;
    if (showit(&a->themap[i][j][k]) != '.') return;
    a->themap[i][j][k].cons = RAND(a) % 4;
    a->themap[i][j][k].s[4] = 1;
    ++a->chest;
}

/* gameplay.hpp:544-557 with gen_zombie / Zombie::gen_npc, Character.hpp:850-871 */
static void spawn_zombie_npc(sfo_arena *a)
{
    int i = RAND(a) % F, j = RAND(a) % N, k = RAND(a) % M;
    if (showit(&a->themap[i][j][k]) != '.') return;
    int index = z_ind(a);
    if (index == -1) return;
    int super_ = (RAND(a) % 4 == 0);
    ozombie *z = &a->zomb[index];
    z->cor[0] = i, z->cor[1] = j, z->cor[2] = k;
    z->super_ = super_;
    z->mindamage = (super_ + 1) * 100;
    z->Hp = (super_ + 1) * 400;
    a->themap[i][j][k].zombie = index;
    a->themap[i][j][k].s[1] = 1;
    a->mz[index] = 1;
}

/* gameplay.hpp:559-572 */
static void spawn_human_npc(sfo_arena *a)
{
    int i = RAND(a) % F, j = RAND(a) % N, k = RAND(a) % M;
    if (showit(&a->themap[i][j][k]) != '.') return;
    int index = h_ind(a);
    if (index == -1) return;
    gen_human(a, 1, &a->hum[index], (int)a->level, i, j, k);
    a->themap[i][j][k].human = index;
    a->themap[i][j][k].s[0] = 1;
    a->mh[index] = 1;
}

/* gameplay.hpp:574-598 */
static void zombie_damage(sfo_arena *a, ocell *pix)
{
    obullet *b = &a->bull[pix->bullet];
    ozombie *z = &a->zomb[pix->zombie];
    pix->s[9] = 1;
    z->Hp -= b->damage; /* Character::hit, Character.hpp:242-246 */
    z->mindamage += b->effect;
    pix->s[2] = 0;
    int owner = b->owner;
    if (owner >= 0) {
        a->hum[owner].damage += b->damage;
        a->hum[owner].effect += b->effect;
    }
    a->mb[pix->bullet] = 0;
    if (z->Hp <= 0) {
        a->mz[pix->zombie] = 0;
        pix->s[8] = 1;
        pix->s[1] = 0;
        if (owner >= 0 && a->hum[owner].team == a->hum[a->ind].team) {
            int pts = 500 + 250 * z->super_;
            ++a->teams_kills, a->loot += pts / 10;
            if (owner == a->ind) a->loot += pts * 9 / 10, ++a->kills;
        }
        if (owner >= 0) ++a->hum[owner].kills;
    }
}

/* gameplay.hpp:600-609 */
static void hit_zombie(sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i)
        if (a->mz[i]) {
            ocell *pix = &a->themap[a->zomb[i].cor[0]][a->zomb[i].cor[1]][a->zomb[i].cor[2]];
            if (pix->s[2]) zombie_damage(a, pix);
        }
}

/* gameplay.hpp:611-634 */
static void human_damage(sfo_arena *a, ocell *pix)
{
    obullet *b = &a->bull[pix->bullet];
    ohuman *h = &a->hum[pix->human];
    pix->s[9] = 1;
    h->Hp -= b->damage;
    h->mindamage += b->effect;
    pix->s[2] = 0;
    int owner = b->owner;
    if (owner >= 0 && h->team != a->hum[owner].team) {
        a->hum[owner].damage += b->damage;
        a->hum[owner].effect += b->effect;
    }
    a->mb[pix->bullet] = 0;
    if (h->Hp <= 0) {
        a->mh[pix->human] = 0;
        pix->s[8] = 1;
        pix->s[0] = (a->ind == pix->human);
        if (owner >= 0 && a->hum[owner].team == a->hum[a->ind].team && h->team != a->hum[a->ind].team) {
            ++a->teams_kills, a->loot += 100;
            if (owner == a->ind) a->loot += 900, ++a->kills;
        }
        if (owner >= 0 && h->team != a->hum[owner].team) ++a->hum[owner].kills;
    }
}

/* gameplay.hpp:636-652 */
static void hit_human(sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i)
        if (a->mh[i]) {
            ocell *pix = &a->themap[a->hum[i].cor[0]][a->hum[i].cor[1]][a->hum[i].cor[2]];
            if (a->hum[i].Hp <= 0) {
                a->mh[i] = 0;
                pix->s[8] = 1;
                pix->s[0] = (pix->human == a->ind);
            } else if (pix->s[2])
                human_damage(a, pix);
            if (a->hum[i].Hp <= 0 && i != a->ind) a->hum[i].agent_active = 0; /* deleteAgent */
        }
}

/* gameplay.hpp:654-693 with Zombie::punch, Character.hpp:838-844 */
static void zombie_action(sfo_arena *a)
{
    for (int zi = 0; zi < MAXS; ++zi)
        if (a->mz[zi]) {
            int i = a->zomb[zi].cor[0], j = a->zomb[zi].cor[1], k = a->zomb[zi].cor[2];
            if (a->themap[i][j][k].s[2]) continue;
            int b = 0;
            for (int i1 = 0; i1 < 4; ++i1) {
                ocell *pix = &a->themap[i][wdx[i1] + j][wdy[i1] + k];
                if (pix->s[0]) {
                    int index = b_ind(a);
                    if (!pix->s[2] && index != -1) {
                        const ozombie *z = &a->zomb[a->themap[i][j][k].zombie];
                        bullet_shot(&a->bull[index], z->cor[0], z->cor[1] + wdx[i1], z->cor[2] + wdy[i1], i1 + 1,
                                    imax(0, z->mindamage), 0, 1, -1);
                        pix->bullet = index;
                        pix->s[2] = 1;
                        a->mb[index] = 1;
                    }
                    b = 1;
                }
            }
            if (!b) {
                if (RAND(a) % 5 < 2) continue;
                for (int i1 = 0; i1 < 2; ++i1) {
                    int i2 = RAND(a) % 4;
                    ocell *t = &a->themap[i][wdx[i2] + j][wdy[i2] + k];
                    if (showit(t) == '.') {
                        t->s[1] = 1;
                        t->zombie = zi;
                        a->themap[i][j][k].s[1] = 0;
                        a->zomb[zi].cor[1] = wdx[i2] + j;
                        a->zomb[zi].cor[2] = wdy[i2] + k;
                        break;
                    }
                }
            }
        }
}

/* gameplay.hpp:695-821 */
static void obey(sfo_arena *a, char c, int hidx)
{
    ohuman *p = &a->hum[hidx];
    if (c == '_') {
        p->Hp = 0;
        return;
    }
    if (c == '[' || c == ']') {
        int d = p->way - 1;
        int f = p->cor[0], r = p->cor[1] + wdx[d], cc = p->cor[2] + wdy[d];
        if (r >= N || 0 > r || cc >= M || 0 > cc) return;
        ocell *x = &a->themap[f][r][cc];
        if (showit(x) != '.') return;
        if (c == '[') {
            if (p->blocks) {
                x->s[10] = x->s[3] = 1;
                --p->blocks;
                a->temp[a->n_temp++] = (f * N + r) * M + cc;
            }
            return;
        }
        if (~p->portal_ind) {
            x->s[10] = x->s[5] = 1;
            x->portal_ind = p->portal_ind;
            p->portal_ind = -1;
            a->temp[a->n_temp++] = (f * N + r) * M + cc;
        } else if (p->portals) {
            int index = p_ind(a);
            if (index == -1) return;
            x->s[10] = x->s[7] = 1;
            --p->portals;
            p->portal_ind = index;
            a->portal[index][0] = f, a->portal[index][1] = r, a->portal[index][2] = cc;
            a->active[index] = 1;
            a->temp[a->n_temp++] = (f * N + r) * M + cc;
        }
        return;
    }
    if (c == 'q' || c == 'e') {
        if (c == 'e') p->way = (p->way == 1) ? 4 : p->way - 1; /* turn_r, Character.hpp:745-751 */
        else p->way = (p->way == 4) ? 1 : p->way + 1;          /* turn_l, :753-759 */
        return;
    }
    if (c == 'a' || c == 's' || c == 'd' || c == 'w') {
        int i = 0;
        const char s[4] = {'s', 'd', 'w', 'a'};
        while (c != s[i]) ++i;
        int f = p->cor[0], r = p->cor[1] + wdx[i], cc = p->cor[2] + wdy[i];
        if (r >= N || 0 > r || cc >= M || 0 > cc) return;
        ocell *t = &a->themap[f][r][cc];
        char sit = showit(t);
        if (sit == '?' || sit == '^' || sit == 'v' || sit == '.' || sit == 'X' || sit == '*') {
            t->s[0] = 1;
            t->human = hidx;
            a->themap[f][p->cor[1]][p->cor[2]].s[0] = 0;
            p->cor[1] = r, p->cor[2] = cc;
        }
        return;
    }
    if (c == 'f' || c == 'g' || c == 'h' || c == 'j') {
        int i = 0;
        const char s[4] = {'f', 'g', 'h', 'j'};
        while (c != s[i]) ++i;
        if (!p->cons[i]) return;
        p->vec = 0, p->ind = i;
        return;
    }
    if (c == 'k' || c == 'l' || c == ';' || c == '\'') {
        int i = 0;
        const char s[4] = {'k', 'l', ';', '\''};
        while (c != s[i]) ++i;
        if (!p->throw_cnt[i]) return;
        p->vec = 1, p->ind = i;
        return;
    }
    if (c == 'c' || c == 'v' || c == 'b' || c == 'n' || c == 'm' || c == ',' || c == '.' || c == '/') {
        int i = 0;
        const char s[8] = {'c', 'v', 'b', 'n', 'm', ',', '.', '/'};
        while (c != s[i]) ++i;
        if (!p->w_lvl[i]) return;
        p->vec = 2, p->ind = i;
        return;
    }
    if (c == 'u') {
        human_use(a, p);
        return;
    }
    if (c == 'z' || c == 'x') {
        int bway = p->way - 1;
        int f = p->cor[0], r = p->cor[1] + wdx[bway], cc = p->cor[2] + wdy[bway];
        int index = b_ind(a);
        if (index == -1 || r >= N || 0 > r || cc >= M || 0 > cc) return;
        int can;
        if (c == 'z') can = human_punch(p, hidx, &a->bull[index]);
        else if (p->vec == 1) can = human_throw_it(a, p, hidx, &a->bull[index]);
        else if (p->vec == 2) can = human_shot_it(a, p, hidx, &a->bull[index]);
        else return;
        ocell *t = &a->themap[f][r][cc];
        char sit = showit(t);
        if (can && ((sit != '#' && sit != 'v' && sit != '^') || t->s[10])) {
            t->bullet = index;
            t->s[2] = 1;
            a->mb[index] = 1;
        }
        return;
    }
}

/* gameplay.hpp:1927-1940 */
static char human_rnpc_bot(sfo_arena *a)
{
    if (a->frame % 50 <= 1) {
        const char c[8] = {'c', 'v', 'b', 'n', 'm', ',', '.', '/'};
        return c[RAND(a) % 8];
    } else if (RAND(a) % 5 < 3)
        return 'x';
    else if (RAND(a) % 5 < 3) {
        const char c[7] = {'1', '2', 'a', 'w', 's', 'd', 'p'};
        return c[RAND(a) % 7];
    }
    const char c[8] = {'+', 'u', 'f', 'g', 'h', 'j', '[', ']'};
    return c[RAND(a) % 8];
}

/* gameplay.hpp:965-1012; agent_cmd[i] is what bot(hum[i]) returns for an agent-driven human
 * (bots/bot-0.5/Custom.hpp:137-158: '+' without an agent, else action[predict(obs)]) */
static void human_action(sfo_arena *a, const uint8_t *agent_cmd, int n_cmd)
{
    for (int i = 0; i < MAXS; ++i)
        if (i != a->ind && a->mh[i]) {
            if (a->remote[i]) a->command[i] = i < n_cmd ? agent_cmd[i] : '+'; /* recieve() / the replay file, :977-986 */
            else if (a->hum[i].rnpc) a->command[i] = (uint8_t)human_rnpc_bot(a);
            else if (a->hum[i].agent_active && i < n_cmd) a->command[i] = agent_cmd[i];
            else a->command[i] = '+';
        }
    int r = RAND(a) & 1, st = (1 - r) * (MAXS - 1), dif = 2 * r - 1;
    for (int i = st; i < MAXS && (~i); i += dif)
        if (a->mh[i]) {
            obey(a, (char)a->command[i], i);
            teleport(a, i);
            claim_chest(a, &a->hum[i]);
            a->command[i] = '+';
        }
}

/* gameplay.hpp:1059-1100.  place[] only lists the cells to copy back; its reversal (:1074)
 * cannot change the outcome because every listed cell is copied from themap1. */
static void update_bull(sfo_arena *a)
{
    static int place[2 * MAXS][3]; /* scratch; the model is single-threaded */
    int cnt = 0;
    for (int b = 0; b < MAXS; ++b)
        if (a->mb[b]) {
            int i = a->bull[b].cor[0], j = a->bull[b].cor[1], k = a->bull[b].cor[2];
            int d = a->bull[b].way - 1;
            a->map1_s2[i][j][k] = 0;
            place[cnt][0] = i, place[cnt][1] = j, place[cnt][2] = k, ++cnt;
            a->map1_s2[i][j + wdx[d]][k + wdy[d]] = 0;
            place[cnt][0] = i, place[cnt][1] = j + wdx[d], place[cnt][2] = k + wdy[d], ++cnt;
        }
    int r = RAND(a) & 1, st = (1 - r) * (MAXS - 1), dif = 2 * r - 1;
    for (int b = st; b < MAXS && (~b); b += dif)
        if (a->mb[b]) {
            obullet *bl = &a->bull[b];
            int i = bl->cor[0], j = bl->cor[1], k = bl->cor[2];
            int dist = abs(bl->cor[0] - bl->dcor[0]) + abs(bl->cor[1] - bl->dcor[1]) + abs(bl->cor[2] - bl->dcor[2]);
            if (dist + 1 >= bl->range) { /* Bullet::expire, Item.hpp:165-168 */
                a->mb[b] = 0;
                continue;
            }
            int d = bl->way - 1;
            const ocell *nx = &a->themap[i][j + wdx[d]][k + wdy[d]];
            char sit = showit(nx);
            if ((sit != '#' && sit != 'v' && sit != '^') || nx->s[10]) {
                a->map1_bullet[i][j + wdx[d]][k + wdy[d]] = b;
                bl->cor[1] = j + wdx[d], bl->cor[2] = k + wdy[d];
                a->map1_s2[i][j + wdx[d]][k + wdy[d]] = 1;
            } else
                a->mb[b] = 0;
        }
    for (int q = 0; q < cnt; ++q) {
        int i = place[q][0], j = place[q][1], k = place[q][2];
        a->themap[i][j][k].s[2] = a->map1_s2[i][j][k];
        a->themap[i][j][k].bullet = a->map1_bullet[i][j][k];
    }
}

/* gameplay.hpp:1279-1297 */
static void portal_damage(sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i) {
        if (!a->active[i]) continue;
        ocell *x = &a->themap[a->portal[i][0]][a->portal[i][1]][a->portal[i][2]];
        if (showit(x) != 'O') {
            int index = b_ind(a);
            if (index == -1) return;
            bullet_shot(&a->bull[index], a->portal[i][0], a->portal[i][1], a->portal[i][2], 3, 20, -10, 1, -1);
            x->bullet = index;
            x->s[2] = 1;
            a->mb[index] = 1;
        }
    }
}

/* gameplay.hpp:1343-1381 */
static void update_tmp(sfo_arena *a)
{
    for (int b = 0; b < MAXS; ++b) {
        if (!a->mb[b]) continue;
        ocell *x = &a->themap[a->bull[b].cor[0]][a->bull[b].cor[1]][a->bull[b].cor[2]];
        char sit = showit(x);
        if ((sit == '^' || sit == '#') && x->s[10]) {
            x->dmg += a->bull[b].damage;
            x->s[9] = 1;
            x->s[2] = 0;
            a->mb[b] = 0;
        }
    }
    ocell *base = &a->themap[0][0][0];
    for (int q = 0; q < a->n_temp; ++q) {
        ocell *e = base + a->temp[q];
        char c = showit(e);
        if (c == '^' && e->dmg >= LIM_PORTAL) {
            int i = e->portal_ind;
            ocell *e1 = &a->themap[a->portal[i][0]][a->portal[i][1]][a->portal[i][2]];
            e1->s[7] = e1->s[10] = 0;
            e->s[5] = e->s[10] = 0;
            e->portal_ind = -1;
            e->dmg = 0;
            a->active[i] = 0;
        } else if (c == '#' && e->dmg >= LIM_BLOCK) {
            e->s[3] = e->s[10] = 0;
            e->dmg = 0;
        }
    }
    for (int q = 0; q < a->n_temp; ++q)
        if (!(base + a->temp[q])->s[10]) {
            a->temp[q] = a->temp[a->n_temp - 1];
            --a->n_temp;
            --q;
        }
}

/* setup(), gameplay.hpp:1231-1277, then the offline placement of load_data(), :1861-1920 */
static void setup(sfo_arena *a)
{
    a->loot = a->teams_kills = a->kills = a->frame = 0;
    a->chest = 0; /* harness: fresh-process semantics (setup() itself never clears it) */
    a->n_temp = 0;
    memset(a->active, 0, sizeof a->active);
    memset(a->mb, 0, sizeof a->mb);
    memset(a->mz, 0, sizeof a->mz);
    memset(a->mh, 0, sizeof a->mh);
    memset(a->command, '+', sizeof a->command);
    for (int i = 0; i < MAXS; ++i) a->hum[i].agent_active = 0;
    memset(a->map1_s2, 0, sizeof a->map1_s2);
    memset(a->map1_bullet, 0xff, sizeof a->map1_bullet);
    for (int k = 0; k < F; ++k)
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < M; ++j) {
                ocell *x = &a->themap[k][i][j];
                int id = (k * N + i) * M + j;
                memset(x, 0, sizeof *x);
                x->portal_ind = -1;
                x->human = x->zombie = x->bullet = x->cons = -1;
                char c = (char)a->map_cells[id];
                if (c == '#') x->s[3] = 1;
                else if (c == '^') x->s[5] = 1, x->portal_ind = a->map_portal[id];
                else if (c == 'v') x->s[6] = 1, x->portal_ind = a->map_portal[id];
                else if (c == 'O') {
                    x->s[7] = 1;
                    int index = p_ind(a);
                    a->portal[index][0] = k, a->portal[index][1] = i, a->portal[index][2] = j;
                    a->active[index] = 1;
                }
            }
    const int me = a->mode == SF_MODE_ROYALE ? a->royale_ind : 0; /* "players ind team" of the match header, :1797-1799 */
    a->ind = me;
    memset(a->remote, 0, sizeof a->remote);
    a->mh[me] = 1;
    memset(&a->hum[me], 0, sizeof a->hum[me]);
    human_build(&a->hum[me], a->player_sheet, 0); /* hum[ind] = me, me.build(false, "", sheet) */
    a->hum[me].way = 1;
    a->hum[me].team = 1;
    if (a->mode == SF_MODE_SQUAD) {
        a->hum[0].cor[0] = 0, a->hum[0].cor[1] = 3, a->hum[0].cor[2] = 1;
        a->themap[0][3][1].human = 0, a->themap[0][3][1].s[0] = 1;
        for (int i = 1; i < 5; ++i) {
            a->mh[i] = 1;
            gen_human(a, 0, &a->hum[i], (int)a->level, 0, 1, i + 1);
            a->themap[0][1][i + 1].human = i, a->themap[0][1][i + 1].s[0] = 1;
            a->hum[i].team = 1;
            a->hum[i].agent_active = a->squad_agents; /* USE_AGENT_IN_SQUAD_NPCS, :1886-1888 */
        }
        for (int i = 5; i < 10; ++i) {
            a->mh[i] = 1;
            gen_human(a, 0, &a->hum[i], (int)a->level, 2, 1, i + 1);
            a->themap[2][1][i + 1].human = i, a->themap[2][1][i + 1].s[0] = 1;
            a->hum[i].team = 2;
            a->hum[i].agent_active = a->squad_agents;
        }
    } else if (a->mode == SF_MODE_ROYALE) {
        /* load_data(), online branch as the replay reader runs it (:1776-1806): every player's
           sheet comes from the match header.  The cells are drawn after the stream is seeded
           (place_players) */
        a->hum[me].team = a->teams[me];
        for (int i = 0; i < a->n_players; ++i) {
            if (i == me) continue;
            memset(&a->hum[i], 0, sizeof a->hum[i]);
            human_build(&a->hum[i], a->royale_sheets[i], 0);
            a->hum[i].team = a->teams[i];
            a->mh[i] = 1;
            a->remote[i] = 1;
            a->hum[i].agent_active = 1; /* not the reference's flag: lets sfo_observe show what that player's own client sees */
        }
    } else {
        a->hum[0].cor[0] = 0, a->hum[0].cor[1] = 1, a->hum[0].cor[2] = 1;
        a->themap[0][1][1].human = 0, a->themap[0][1][1].s[0] = 1;
    }
    a->hum[me].agent_active = 1; /* prepare(me), :1743-1744 */
}

/* gameplay.hpp:1847-1859: way and a rejection-sampled '.' cell for every player, in index order */
static void place_players(sfo_arena *a)
{
    for (int i = 0; i < a->n_players; ++i) {
        a->hum[i].way = RAND(a) % 4 + 1;
        for (;;) {
            int f = RAND(a) % F, r = RAND(a) % N, c = RAND(a) % M;
            if (showit(&a->themap[f][r][c]) == '.') {
                a->themap[f][r][c].human = i;
                a->themap[f][r][c].s[0] = 1;
                a->hum[i].cor[0] = f, a->hum[i].cor[1] = r, a->hum[i].cor[2] = c;
                break;
            }
        }
    }
}

/* harness track_and_check_caps() */
static void track_and_check_caps(sfo_arena *a)
{
    int hi_h = -1, hi_z = -1, hi_b = -1, hi_p = -1;
    for (int i = 0; i < MAXS; ++i) {
        if (a->mh[i]) hi_h = i;
        if (a->mz[i]) hi_z = i;
        if (a->mb[i]) hi_b = i;
        if (a->active[i]) hi_p = i;
    }
    if (hi_h + 1 > a->hw_h) a->hw_h = hi_h + 1;
    if (hi_h >= a->cap_h || hi_z >= a->cap_z || hi_b >= a->cap_b || hi_p >= a->cap_portal ||
        a->chest > a->cap_chest || a->n_temp > a->cap_built)
        a->status = SF_OVERFLOW;
}

/* harness update_bull_would_go_out_of_bounds(): gameplay.hpp:1069 reads the cell in front of
 * every live bullet without a bounds test */
static int update_bull_oob(const sfo_arena *a)
{
    for (int i = 0; i < MAXS; ++i)
        if (a->mb[i]) {
            int d = a->bull[i].way - 1;
            int r = a->bull[i].cor[1] + wdx[d], c = a->bull[i].cor[2] + wdy[d];
            if (r < 0 || r >= N || c < 0 || c >= M) return 1;
        }
    return 0;
}

/* harness eval_end(): check_end(), gameplay.hpp:1102-1229, offline branches, frame clock */
static void eval_end(sfo_arena *a)
{
    if (a->status != SF_RUNNING) return;
    if (a->mode == SF_MODE_ROYALE && rivals_are_dead(a)) { /* "online && rivals_are_dead()" comes first, :1103 */
        a->status = SF_WIN;
        return;
    }
    if (a->hum[a->ind].Hp <= 0) {
        a->status = SF_DEAD;
        return;
    }
    if (a->mode == SF_MODE_ROYALE) {
        /* no other way to end an online match */
    } else if (a->mode == SF_MODE_TIMER) {
        if (a->frame >= a->level * 7500) a->status = (a->kills < a->level * 5) ? SF_TIMEOUT : SF_WIN;
    } else if (a->mode == SF_MODE_SOLO) {
        if (a->level * 5 <= a->kills) a->status = SF_WIN;
    } else {
        if (a->level * 10 <= a->teams_kills && rivals_are_dead(a)) a->status = SF_WIN;
    }
    if (a->status == SF_RUNNING && a->max_steps > 0 && a->steps >= a->max_steps) a->status = SF_TRUNCATED;
}

/* ------------------------------------------------------------------ public API */

sfo_arena *sfo_create(const sf_config *cfg)
{
    if (!cfg || !cfg->map_cells || !cfg->map_portal) return NULL;
    if (cfg->cap_humans >= MAXS / 2 || cfg->cap_zombies >= MAXS / 2 || cfg->cap_bullets >= MAXS / 2 ||
        cfg->cap_portals >= MAXS / 2)
        return NULL;
    sfo_arena *a = (sfo_arena *)calloc(1, sizeof *a);
    if (!a) return NULL;
    a->mode = cfg->mode;
    if (cfg->mode == SF_MODE_ROYALE) {
        if (cfg->royale_players < 2 || cfg->royale_players > SF_MAX_PLAYERS) return free(a), (sfo_arena *)NULL;
        a->n_players = cfg->royale_players;
        for (int i = 0; i < a->n_players; ++i) a->teams[i] = cfg->royale_teams[i];
        memcpy(a->royale_sheets, cfg->royale_sheets, sizeof a->royale_sheets);
        if (cfg->royale_ind < 0 || cfg->royale_ind >= a->n_players) return free(a), (sfo_arena *)NULL;
        a->royale_ind = cfg->royale_ind;
    }
    a->squad_agents = cfg->squad_agents != 0;
    a->max_steps = cfg->max_steps;
    a->cap_h = cfg->cap_humans, a->cap_z = cfg->cap_zombies, a->cap_b = cfg->cap_bullets;
    a->cap_chest = cfg->cap_chests, a->cap_built = cfg->cap_built, a->cap_portal = cfg->cap_portals;
    memcpy(a->map_cells, cfg->map_cells, SF_CELLS);
    memcpy(a->map_portal, cfg->map_portal, SF_CELLS * sizeof(int16_t));
    memcpy(a->cons, cfg->consumables, sizeof a->cons);
    memcpy(a->thr, cfg->throwables, sizeof a->thr);
    memcpy(a->wpn, cfg->weapons, sizeof a->wpn);
    memcpy(a->player_sheet, cfg->player_sheet, sizeof a->player_sheet);
    memcpy(a->npc_sheet, cfg->npc_sheet, sizeof a->npc_sheet);
    a->level = cfg->level_min > 0 ? cfg->level_min : 1;
    sfo_reset(a, (int)a->level, sf_synth_tb(0), sf_synth_serial(0, 0));
    return a;
}

void sfo_destroy(sfo_arena *a) { free(a); }

void sfo_reset(sfo_arena *a, int level, int64_t tb, int64_t serial)
{
    a->level = level;
    if (a->mode == SF_MODE_ROYALE) level = 1; /* gameplay.hpp:1641, 1659 */
    a->level = level;
    setup(a);
    sfo_srand(&a->rng, tb, serial);
    a->jomle0 = a->rng.jomle;
    if (a->mode == SF_MODE_ROYALE) place_players(a);
    ++a->frame; /* play(): "++frame" before the loop, gameplay.hpp:1441 */
    a->status = SF_RUNNING;
    a->steps = 0;
    a->hw_h = 0;
    memset(&a->out, 0, sizeof a->out);
    track_and_check_caps(a);
}

int sfo_status(const sfo_arena *a) { return a->status; }

#define CHECK(a) do { track_and_check_caps(a); if ((a)->status != SF_RUNNING) return (a)->status; } while (0)
#define UBGUARD(a) do { if (update_bull_oob(a)) { (a)->status = SF_UB_GUARD; return (a)->status; } } while (0)

typedef struct snap { int64_t kills, teams_kills, loot; int hp, damage, effect; } snap;
static snap take_snap(const sfo_arena *a)
{
    snap s = {a->kills, a->teams_kills, a->loot, a->hum[a->ind].Hp, a->hum[a->ind].damage, a->hum[a->ind].effect};
    return s;
}
static void add_delta(sfo_arena *a, snap s0)
{
    snap s1 = take_snap(a);
    a->out.d_kills += (int)(s1.kills - s0.kills);
    a->out.d_teams_kills += (int)(s1.teams_kills - s0.teams_kills);
    a->out.d_loot += (int)(s1.loot - s0.loot);
    a->out.d_hp += s1.hp - s0.hp;
    a->out.d_damage += s1.damage - s0.damage;
    a->out.d_effect += s1.effect - s0.effect;
    a->out.status = a->status;
    a->out.episode_steps = (int)a->steps;
}

/* first half of the loop body, gameplay.hpp:1444-1461 */
static int step_a_(sfo_arena *a)
{
    if (a->frame % PC <= 1) spawn_chest(a);
    if (a->frame % PZ <= 1) spawn_zombie_npc(a);
    if (a->frame % PH <= 1) spawn_human_npc(a);
    CHECK(a);
    zombie_action(a);
    CHECK(a);
    portal_damage(a);
    CHECK(a);
    update_tmp(a);
    hit_human(a), hit_zombie(a);
    ++a->frame;
    updmap(a);
    UBGUARD(a);
    update_bull(a);
    return a->status;
}

/* second half, gameplay.hpp:1462-1471, then the harness's end-of-step victory test */
static int step_b_(sfo_arena *a, const uint8_t *actions, int n)
{
    a->command[a->ind] = n > a->ind ? actions[a->ind] : '+'; /* actions[] is indexed by human slot */
    uint8_t cmd[64];
    int nc = n < 64 ? n : 64;
    for (int i = 0; i < nc; ++i) { /* harness action_index(): symbols outside gameplay::action read '+' */
        cmd[i] = strchr(SF_ACTIONS9, actions[i]) && actions[i] ? actions[i] : '+';
        if (a->remote[i]) cmd[i] = actions[i]; /* a remote player's command is taken as sent */
    }
    human_action(a, cmd, nc);
    CHECK(a);
    update_tmp(a);
    hit_human(a), hit_zombie(a);
    ++a->frame;
    updmap(a);
    UBGUARD(a);
    update_bull(a);
    ++a->steps;
    eval_end(a);
    return a->status;
}

int sfo_step_a(sfo_arena *a)
{
    if (a->status != SF_RUNNING) return a->status;
    memset(&a->out, 0, sizeof a->out);
    snap s0 = take_snap(a);
    step_a_(a);
    add_delta(a, s0);
    return a->status;
}

int sfo_step_b(sfo_arena *a, const uint8_t *actions, int n)
{
    if (a->status != SF_RUNNING) return a->status;
    snap s0 = take_snap(a);
    step_b_(a, actions, n);
    add_delta(a, s0);
    return a->status;
}

int sfo_step(sfo_arena *a, const uint8_t *actions, int n)
{
    if (a->status != SF_RUNNING) return a->status;
    memset(&a->out, 0, sizeof a->out);
    snap s0 = take_snap(a);
    if (step_a_(a) == SF_RUNNING) step_b_(a, actions, n);
    add_delta(a, s0);
    return a->status;
}

void sfo_step_out(const sfo_arena *a, sf_step_out *out) { *out = a->out; }
int64_t sfo_rng_draws(const sfo_arena *a) { return a->rng.jomle - a->jomle0; }

void sfo_counters(const sfo_arena *a, int64_t out[8])
{
    out[0] = a->frame, out[1] = a->kills, out[2] = a->teams_kills, out[3] = a->loot;
    out[4] = a->chest, out[5] = a->steps, out[6] = a->status, out[7] = a->hum[a->ind].Hp;
}

void sfo_population(const sfo_arena *a, int32_t out[6])
{
    int nh = 0, nz = 0, nb = 0, np = 0;
    for (int i = 0; i < MAXS; ++i) nh += a->mh[i], nz += a->mz[i], nb += a->mb[i], np += a->active[i];
    out[0] = nh, out[1] = nz, out[2] = nb, out[3] = (int)a->chest, out[4] = a->n_temp, out[5] = np;
}

static int emit(int32_t *buf, long cap, long *n, int kind, int index, const int32_t *f, int nf)
{
    if (*n + 3 + nf > cap) return -1;
    buf[(*n)++] = kind, buf[(*n)++] = index, buf[(*n)++] = nf;
    for (int i = 0; i < nf; ++i) buf[(*n)++] = f[i];
    return 0;
}

/* canonical record, include/sf_canon.h (same element order as harness sfref_dump) */
long sfo_dump(const sfo_arena *a, int32_t *buf, long cap)
{
    long n = 0;
    int32_t f[32];
    f[0] = a->mode, f[1] = (int)a->level, f[2] = (int)a->frame, f[3] = (int)a->kills, f[4] = (int)a->teams_kills;
    f[5] = (int)a->loot, f[6] = (int)a->chest, f[7] = a->ind;
    if (emit(buf, cap, &n, SF_K_HEADER, 0, f, SF_NF_HEADER)) return -1;
    for (int i = 0; i < 18; ++i) f[i] = (int)a->rng.random[i];
    f[18] = (int)(a->rng.jomle & 0xFFFF);
    if (emit(buf, cap, &n, SF_K_RNG, 0, f, SF_NF_RNG)) return -1;
    for (int i = 0; i < a->hw_h; ++i) {
        const ohuman *h = &a->hum[i];
        int k = 0;
        f[k++] = a->mh[i], f[k++] = h->rnpc, f[k++] = h->team, f[k++] = h->way;
        f[k++] = h->cor[0], f[k++] = h->cor[1], f[k++] = h->cor[2];
        f[k++] = h->Hp, f[k++] = h->mindamage, f[k++] = h->stamina;
        f[k++] = h->kills, f[k++] = h->damage, f[k++] = h->effect;
        f[k++] = h->vec, f[k++] = h->ind;
        for (int j = 0; j < 4; ++j) f[k++] = h->cons[j];
        for (int j = 0; j < 4; ++j) f[k++] = h->throw_cnt[j];
        f[k++] = h->blocks, f[k++] = h->portals, f[k++] = h->portal_ind;
        f[k++] = h->mindamage_def;
        if (emit(buf, cap, &n, SF_K_HUMAN, i, f, SF_NF_HUMAN)) return -1;
    }
    for (int i = 0; i < MAXS; ++i)
        if (a->mz[i]) {
            const ozombie *z = &a->zomb[i];
            f[0] = z->super_, f[1] = z->cor[0], f[2] = z->cor[1], f[3] = z->cor[2], f[4] = z->Hp, f[5] = z->mindamage;
            if (emit(buf, cap, &n, SF_K_ZOMBIE, i, f, SF_NF_ZOMBIE)) return -1;
        }
    for (int i = 0; i < MAXS; ++i)
        if (a->mb[i]) {
            const obullet *b = &a->bull[i];
            f[0] = b->cor[0], f[1] = b->cor[1], f[2] = b->cor[2], f[3] = b->dcor[0], f[4] = b->dcor[1], f[5] = b->dcor[2];
            f[6] = b->way, f[7] = b->range, f[8] = b->damage, f[9] = b->effect, f[10] = b->owner;
            if (emit(buf, cap, &n, SF_K_BULLET, i, f, SF_NF_BULLET)) return -1;
        }
    for (int i = 0; i < MAXS; ++i)
        if (a->active[i]) {
            f[0] = a->portal[i][0], f[1] = a->portal[i][1], f[2] = a->portal[i][2];
            if (emit(buf, cap, &n, SF_K_PORTAL, i, f, SF_NF_PORTAL)) return -1;
        }
    for (int fl = 0; fl < F; ++fl)
        for (int r = 0; r < N; ++r)
            for (int c = 0; c < M; ++c) {
                const ocell *x = &a->themap[fl][r][c];
                if (!(x->s[0] || x->s[1] || x->s[2] || x->s[4] || x->s[10])) continue;
                int kind = x->s[10] ? (x->s[3] ? 1 : x->s[5] ? 2 : x->s[7] ? 3 : 0) : 0;
                f[0] = x->s[0], f[1] = x->s[0] ? x->human : -1;
                f[2] = x->s[1], f[3] = x->s[1] ? x->zombie : -1;
                f[4] = x->s[2], f[5] = x->s[2] ? x->bullet : -1;
                f[6] = x->s[4], f[7] = x->s[4] ? x->cons : -1;
                f[8] = kind, f[9] = x->dmg, f[10] = kind == 2 ? x->portal_ind : -1;
                if (emit(buf, cap, &n, SF_K_CELL, (fl * N + r) * M + c, f, SF_NF_CELL)) return -1;
            }
    return n;
}

uint64_t sfo_hash(const sfo_arena *a)
{
    static int32_t buf[1 << 20];
    long n = sfo_dump(a, buf, 1 << 20);
    return n < 0 ? 0ULL : sf_canon_hash(buf, n);
}

/* ------------------------------------------------------------------ bots/bot-0.5/Custom.hpp */

/* describe(), Custom.hpp:29-135: 32 floats for one cell as seen by `player` */
static void describe(const sfo_arena *a, const ocell *cell, const ohuman *player, float res[32])
{
    int k = 0;
    res[k++] = (float)(cell->s[0] || cell->s[1]);
    res[k++] = (float)cell->s[2];
    res[k++] = (float)cell->s[3];
    res[k++] = (float)cell->s[4];
    res[k++] = (float)(cell->s[5] || cell->s[6]);
    res[k++] = (float)cell->s[7];
    res[k++] = (float)cell->s[10];
    float sit[4] = {0, 0, 0, 0};
    if (cell->s[0]) {
        int t = a->hum[cell->human].team;
        if (!t) sit[2] = 1;
        else if (t == player->team) sit[0] = 1;
        else sit[1] = 1;
    }
    if (cell->s[1]) sit[3] = 1;
    for (int i = 0; i < 4; ++i) res[k++] = sit[i];
    if (cell->s[0]) {
        const ohuman *h = &a->hum[cell->human];
        res[k++] = (float)h->kills;
        res[k++] = (float)h->blocks;
        res[k++] = (float)h->portals;
        res[k++] = (float)(h->portal_ind != -1);
    } else
        for (int i = 0; i < 4; ++i) res[k++] = 0;
    sit[0] = sit[1] = sit[2] = 0;
    float hp = 0;
    if (cell->s[3] || cell->s[5] || cell->s[6] || cell->s[0] || cell->s[1]) {
        sit[0] = sit[1] = 1;
        sit[2] = (float)(cell->s[10] || cell->s[0] || cell->s[1]);
        if (cell->s[0]) hp = (float)(a->hum[cell->human].Hp / 1000.0);
        else if (cell->s[1]) hp = (float)(a->zomb[cell->zombie].Hp / 1000.0);
        else if (cell->s[10]) {
            if (cell->s[3]) hp = (float)((LIM_BLOCK - cell->dmg) / 1000.0);
            else hp = (float)((LIM_PORTAL - cell->dmg) / 1000.0);
        }
    } else if (cell->s[7]) {
        sit[0] = 1;
        sit[1] = sit[2] = 0;
    }
    for (int i = 0; i < 3; ++i) res[k++] = sit[i];
    res[k++] = hp;
    sit[0] = sit[1] = sit[2] = sit[3] = 0;
    float damage = 0, effect = 0, is_bull = 0, estamina = 0;
    if (cell->s[0]) {
        const ohuman *h = &a->hum[cell->human];
        int v[2];
        sit[h->way - 1] = 1;
        human_damage_effect(a, h, v);
        damage = (float)(v[0] / 1000.0);
        effect = (float)(-v[1] / 1000.0);
        estamina = (float)(h->stamina / 1000.0);
    } else if (cell->s[1]) {
        sit[0] = sit[1] = sit[2] = sit[3] = (float)0.01;
        damage = (float)(a->zomb[cell->zombie].mindamage / 1000.0);
    } else if (cell->s[2]) {
        const obullet *b = &a->bull[cell->bullet];
        is_bull = 1;
        int dist_traveled = abs(b->cor[1] - b->dcor[1]) + abs(b->cor[2] - b->dcor[2]);
        sit[b->way - 1] = (float)((b->range - dist_traveled) / 100.0);
        damage = (float)(b->damage / 1000.0);
        effect = (float)(-b->effect / 1000.0);
    } else if (cell->s[7]) {
        damage = (float)(20 / 1000.0);
        effect = (float)(10 / 1000.0);
    }
    res[k++] = is_bull;
    for (int i = 0; i < 4; ++i) res[k++] = sit[i];
    res[k++] = damage, res[k++] = effect, res[k++] = estamina;
    sit[0] = sit[1] = sit[2] = 0;
    if (cell->s[4]) {
        sit[0] = (float)(a->cons[cell->cons].stamina / 1000.0);
        sit[1] = (float)(a->cons[cell->cons].effect / 1000.0);
        sit[2] = (float)(a->cons[cell->cons].hp / 1000.0);
    }
    for (int i = 0; i < 3; ++i) res[k++] = sit[i];
    if (cell->s[0]) {
        res[k++] = (float)(a->hum[cell->human].damage / 1000.0);
        res[k++] = (float)(-a->hum[cell->human].effect / 1000.0);
    } else {
        res[k++] = 0;
        res[k++] = 0;
    }
}

/* Custom.hpp:157: obs.push_back(std::pow(std::abs(ch) / 10, 0.2)) with ch a float:
 * float abs, float division by 10, then pow in double, rounded to float by the vector<float> */
float sfo_obs_transform(float x) { return (float)pow((double)(fabsf(x) / 10), 0.2); }

static int observe_(const sfo_arena *a, int slot, float *out, int raw)
{
    if (slot < 0 || slot >= MAXS || !a->hum[slot].agent_active) return -1;
    const ohuman *p = &a->hum[slot];
    ocell nd;
    memset(&nd, 0, sizeof nd);
    nd.portal_ind = -1, nd.human = nd.zombie = nd.bullet = nd.cons = -1;
    const int R = SF_OBS_WIN / 2;
    float vec[32];
    for (int i = p->cor[1] - R, wi = 0; i <= p->cor[1] + R; ++i, ++wi)
        for (int j = p->cor[2] - R, wj = 0; j <= p->cor[2] + R; ++j, ++wj) {
            if (i < 0 || j < 0 || N <= i || M <= j) describe(a, &nd, p, vec);
            else describe(a, &a->themap[p->cor[0]][i][j], p, vec);
            for (int k = 0; k < 32; ++k)
                out[k * SF_OBS_WIN * SF_OBS_WIN + wi * SF_OBS_WIN + wj] = raw ? vec[k] : sfo_obs_transform(vec[k]);
        }
    return SF_OBS_LEN;
}

int sfo_observe(const sfo_arena *a, int slot, float *out) { return observe_(a, slot, out, 0); }
int sfo_observe_raw(const sfo_arena *a, int slot, float *out) { return observe_(a, slot, out, 1); }

/* ------------------------------------------------------------------ synthetic workload */

long sfo_run_stream(sfo_arena *a, int64_t env, int level, const char *table, int table_len, long n_steps,
                    int with_obs, uint64_t *hash_out)
{
    static float obs[SF_OBS_LEN];
    int64_t episode = 0;
    uint64_t streams[SF_MAX_PLAYERS];
    for (int ag = 0; ag < SF_MAX_PLAYERS; ++ag) streams[ag] = sf_synth_stream_init(env, ag);
    int n_agents = a->mode == SF_MODE_ROYALE ? a->n_players : (a->mode == SF_MODE_SQUAD && a->squad_agents) ? 10 : 1;
    sfo_reset(a, level, sf_synth_tb(env), sf_synth_serial(env, episode));
    uint8_t act[SF_MAX_PLAYERS];
    uint64_t acc = 0;
    for (long s = 0; s < n_steps; ++s) {
        for (int ag = 0; ag < SF_MAX_PLAYERS; ++ag) {
            uint64_t z = sf_synth_stream_next(&streams[ag]);
            act[ag] = (uint8_t)table[z % (uint64_t)table_len];
        }
        if (with_obs) {
            sfo_observe(a, 0, obs);
            acc += (uint64_t)(obs[15 * 31 + 15] * 1000.f);
        }
        if (sfo_step(a, act, n_agents) != SF_RUNNING) {
            ++episode;
            sfo_reset(a, level, sf_synth_tb(env), sf_synth_serial(env, episode));
        }
    }
    if (hash_out) *hash_out = sfo_hash(a) + acc;
    return n_steps;
}

/* The same free-running workload, but with a CHECKSUM OF THE WHOLE TRAJECTORY: after every step
 * (and any auto-reset it triggers) the canonical-state hash is folded into
 * chk = mix64(chk ^ hash).  Optionally the observation of human slot 0 after the last step.
 * This is what the full-size GPU parity tests compare against (one value per arena covers every
 * step of it).  Returns the number of episodes that ended. */
long sfo_run_trace(sfo_arena *a, int64_t env, int level, const char *table, int table_len, long n_steps,
                   uint64_t *chk_out, uint64_t *last_hash_out, float *obs_out)
{
    int64_t episode = 0;
    uint64_t streams[SF_MAX_PLAYERS];
    for (int ag = 0; ag < SF_MAX_PLAYERS; ++ag) streams[ag] = sf_synth_stream_init(env, ag);
    int n_agents = a->mode == SF_MODE_ROYALE ? a->n_players : (a->mode == SF_MODE_SQUAD && a->squad_agents) ? 10 : 1;
    sfo_reset(a, level, sf_synth_tb(env), sf_synth_serial(env, episode));
    uint8_t act[SF_MAX_PLAYERS];
    uint64_t chk = 0, h = sfo_hash(a);
    for (long s = 0; s < n_steps; ++s) {
        for (int ag = 0; ag < SF_MAX_PLAYERS; ++ag) {
            uint64_t z = sf_synth_stream_next(&streams[ag]);
            act[ag] = (uint8_t)table[z % (uint64_t)table_len];
        }
        if (sfo_step(a, act, n_agents) != SF_RUNNING) {
            ++episode;
            sfo_reset(a, level, sf_synth_tb(env), sf_synth_serial(env, episode));
        }
        h = sfo_hash(a);
        chk = sf_mix64(chk ^ h);
    }
    if (chk_out) *chk_out = chk;
    if (last_hash_out) *last_hash_out = h;
    if (obs_out && sfo_observe(a, 0, obs_out) != SF_OBS_LEN) obs_out[0] = -1.0f; /* no active agent: the reference builds nothing */
    return (long)episode;
}
