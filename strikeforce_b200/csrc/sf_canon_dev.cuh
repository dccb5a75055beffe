/*
 * sf_canon_dev.cuh -- the canonical record (include/sf_canon.h) of one arena, produced from
 * the device layout.  The element order and the fields are those of the reference harness's
 * sfref_dump, so a record exported here can be compared int for int with one taken from the
 * unmodified reference.  Used by the export / state-hash kernels (parity checks); not on the
 * step path.
 */
#ifndef SF_CANON_DEV_CUH
#define SF_CANON_DEV_CUH

#include "sf_canon.h"
#include "sf_core.cuh"

/* sinks: a flat int32 buffer, or the order-independent hash */
struct SfBufSink {
    static constexpr bool ordered = true; /* elements in record order (cells by index) */
    int32_t *buf;
    long cap, n;
    bool overflow;
    SF_MFN void elem(int kind, int index, const int32_t *f, int nf)
    {
        if (n + 3 + nf > cap) {
            overflow = true;
            return;
        }
        buf[n++] = kind, buf[n++] = index, buf[n++] = nf;
        for (int i = 0; i < nf; ++i) buf[n++] = f[i];
    }
};
struct SfHashSink {
    static constexpr bool ordered = false; /* a sum: any order */
    uint64_t sum;
    SF_MFN void elem(int kind, int index, const int32_t *f, int nf)
    {
        uint64_t h = sf_canon_elem_begin(kind, index);
        for (int i = 0; i < nf; ++i) h = sf_canon_elem_field(h, f[i]);
        sum += h;
    }
};

/* one SF_K_CELL element */
template <class Sink>
SF_FN void sf_canon_cell(const SfDev &d, int env, const SfEnv &e, int cell, int lin, uint32_t g, int bidx, Sink &sink)
{
    int32_t f[SF_NF_CELL];
    uint32_t kind = (g >> C_KIND_SHIFT) & 7u;
    bool chest = kind >= K_CHEST0 && kind < K_BLOCK;
    int built = kind == K_BLOCK ? 1 : kind == K_ENTRANCE ? 2 : kind == K_EXIT ? 3 : 0;
    int q = built ? sf_find_built(d, env, e, cell) : -1;
    f[0] = (g & C_S0) ? 1 : 0, f[1] = (g & C_S0) ? (int32_t)(g & C_OCC) : -1;
    f[2] = (g & C_S1) ? 1 : 0, f[3] = (g & C_S1) ? (int32_t)(g & C_OCC) : -1;
    f[4] = bidx >= 0 ? 1 : 0, f[5] = bidx;
    f[6] = chest ? 1 : 0, f[7] = chest ? (int32_t)kind - K_CHEST0 : -1;
    f[8] = built, f[9] = q >= 0 ? SF_T(d.t_dmg, q) : 0, f[10] = built == 2 ? (int32_t)SF_T(d.t_pidx, q) : -1;
    sink.elem(SF_K_CELL, lin, f, SF_NF_CELL);
}

template <class Sink>
SF_FN void sf_canon_emit(const SfDev &d, const SfConst &k, const SfTabs &t, int env, Sink &sink)
{
    SfEnv e;
    sf_load_env(d, env, e);
    int32_t f[32];
    f[0] = k.mode, f[1] = e.level, f[2] = (int32_t)e.frame, f[3] = e.kills, f[4] = e.tkills, f[5] = e.loot;
    f[6] = e.chest, f[7] = k.ind;
    sink.elem(SF_K_HEADER, 0, f, SF_NF_HEADER);
    for (int i = 0; i < 18; ++i) f[i] = (int32_t)sf_exp_m1(t, 2u * sf_rng_log(e, i)) + 1;
    f[18] = (int32_t)(e.jomle & 0xFFFFu);
    sink.elem(SF_K_RNG, 0, f, SF_NF_RNG);
    for (int h = 0; h < e.hw_h; ++h) {
        uint32_t pw = SF_AT(d.h_pw, h), sel = SF_AT(d.h_sel, h), bp = SF_AT(d.h_bp, h);
        uint32_t cp = SF_AT(d.h_cons, h), tp = SF_AT(d.h_thr, h);
        int n = 0, pf, pr, pc;
        sf_tcell_decode((int)(pw & POS_CELL), &pf, &pr, &pc);
        f[n++] = (int32_t)((e.mh >> h) & 1), f[n++] = (sel & HS_RNPC) ? 1 : 0, f[n++] = sf_team_value(sel);
        f[n++] = (int32_t)(pw >> POS_HI_SHIFT) + 1;
        f[n++] = pf, f[n++] = pr, f[n++] = pc;
        f[n++] = SF_AT(d.h_hp, h), f[n++] = SF_AT(d.h_mind, h), f[n++] = SF_AT(d.h_stam, h);
        f[n++] = SF_AT(d.h_kills, h), f[n++] = SF_AT(d.h_dmg, h), f[n++] = SF_AT(d.h_eff, h);
        f[n++] = (int32_t)((sel >> HS_VEC_SHIFT) & 3u) - 1, f[n++] = (int32_t)((sel >> HS_IND_SHIFT) & 15u) - 1;
        for (int j = 0; j < 4; ++j) f[n++] = (int32_t)((cp >> (8 * j)) & 0xFFu);
        for (int j = 0; j < 4; ++j) f[n++] = (int32_t)((tp >> (8 * j)) & 0xFFu);
        f[n++] = (int32_t)(bp & 0xFFu), f[n++] = (int32_t)((bp >> 8) & 0xFFu), f[n++] = (int32_t)((bp >> 16) & 0xFFu) - 1;
        f[n++] = h < k.n_players ? k.players[h].mindamage_def : k.npc_mindamage_def[e.level];
        sink.elem(SF_K_HUMAN, h, f, SF_NF_HUMAN);
    }
    for (int z = m2_next(e.mz, 0); z >= 0; z = m2_next(e.mz, z + 1)) {
        uint32_t pw = SF_AT(d.z_pos, z);
        int pf, pr, pc;
        sf_tcell_decode((int)(pw & POS_CELL), &pf, &pr, &pc);
        f[0] = (int32_t)(pw >> POS_HI_SHIFT), f[1] = pf, f[2] = pr;
        f[3] = pc, f[4] = SF_AT(d.z_hp, z), f[5] = SF_AT(d.z_mind, z);
        sink.elem(SF_K_ZOMBIE, z, f, SF_NF_ZOMBIE);
    }
    for (int b = m2_next(e.mb, 0); b >= 0; b = m2_next(e.mb, b + 1)) {
        uint32_t pw = SF_AT(d.b_pw, b), meta = SF_AT(d.b_meta, b);
        int way0 = (int)(pw >> POS_HI_SHIFT), trav = (int)((meta >> 8) & 0xFFu);
        int fl, r, c;
        sf_tcell_decode((int)(pw & POS_CELL), &fl, &r, &c);
        f[0] = fl, f[1] = r, f[2] = c;
        f[3] = fl, f[4] = r - trav * ((way0 == 0) - (way0 == 2)), f[5] = c - trav * ((way0 == 1) - (way0 == 3));
        f[6] = way0 + 1, f[7] = (int32_t)(meta & 0xFFu), f[8] = SF_AT(d.b_dmg, b), f[9] = SF_AT(d.b_eff, b);
        f[10] = (int32_t)((meta >> 16) & 0xFFu) - 1;
        sink.elem(SF_K_BULLET, b, f, SF_NF_BULLET);
    }
    for (int p = m2_next(e.mp, 0); p >= 0; p = m2_next(e.mp, p + 1)) {
        sf_tcell_decode(sf_exit_cell(d, k, env, p), &f[0], &f[1], &f[2]);
        sink.elem(SF_K_PORTAL, p, f, SF_NF_PORTAL);
    }
    /* the bullet flags of the record (s2, bidx) come from the owning bullets; their cells first */
    uint16_t oc[SF_LIM_BULLETS];
    int16_t ob[SF_LIM_BULLETS];
    int n_own = 0;
    for (int b = m2_next(e.mb, 0); b >= 0; b = m2_next(e.mb, b + 1))
        if (SF_AT(d.b_meta, b) & BF_OWNS) oc[n_own] = (uint16_t)(SF_AT(d.b_pw, b) & POS_CELL), ob[n_own++] = (int16_t)b;
    for (int lin = 0; lin < SF_CELLS; ++lin) { /* the record lists cells in the reference's (floor, row, col) order */
        const int cell = sf_tcell(lin / (SF_ROWS * SF_COLS), (lin / SF_COLS) % SF_ROWS, lin % SF_COLS);
        uint32_t g = SF_G(cell);
        int bidx = -1;
        if (Sink::ordered || g)
            for (int i = 0; i < n_own; ++i)
                if (oc[i] == cell) bidx = ob[i];
        uint32_t kind = (g >> C_KIND_SHIFT) & 7u;
        if (!Sink::ordered && !g) continue; /* cells that carry a bullet flag only follow below */
        if (!(g & (C_S0 | C_S1)) && bidx < 0 && kind == K_NONE) continue;
        sf_canon_cell(d, env, e, cell, lin, g, bidx, sink);
    }
    if (!Sink::ordered)
        for (int i = 0; i < n_own; ++i)
            if (!SF_G(oc[i])) {
                int f0, r0, c0;
                sf_tcell_decode(oc[i], &f0, &r0, &c0);
                sf_canon_cell(d, env, e, oc[i], (f0 * SF_ROWS + r0) * SF_COLS + c0, 0u, ob[i], sink);
            }
}

#endif
