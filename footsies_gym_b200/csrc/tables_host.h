// tables_host.h -- fills fg::Tables on the host from the generated frame data (frame_tables.h) and the BattleAI
// input sequences (BattleAI.cs:192-342).  Included by the CUDA library (uploaded once per handle) and by the
// host-emulation test harness.
#ifndef FOOTSIES_B200_TABLES_HOST_H
#define FOOTSIES_B200_TABLES_HOST_H

#include <string.h>

#include <vector>

#include "frame_logic.cuh"

namespace fg {

inline void build_tables(Tables &t) {
    memset(&t, 0, sizeof t);
    static const uint32_t rows[FT_NUM_ROWS][4] = FT_ROWS_INIT;
    static const uint32_t boxcfg[FT_NUM_BOXCFG][8] = FT_BOXCFG_INIT;
    static const uint32_t attack[5][8] = FT_ATTACK_INIT;
    static const uint8_t cum_next[FT_NUM_CUM][4] = FT_CUM_NEXT_INIT;
    static const double step_reward[4] = FT_STEP_REWARD_INIT;
    static const double term[FT_NUM_CUM][4][2] = FT_TERM_REWARD_INIT;
    for (int i = 0; i < FT_NUM_ROWS; i++) { t.rows[i].x = rows[i][0]; t.rows[i].y = rows[i][1]; t.rows[i].z = rows[i][2]; t.rows[i].w = rows[i][3]; }
    static_assert(sizeof(BoxCfg) == sizeof(boxcfg[0]) && sizeof(AttackRow) == sizeof(attack[0]), "generated row size");
    memcpy(t.boxcfg, boxcfg, sizeof boxcfg);
    memcpy(t.attack, attack, sizeof attack);
    memcpy(t.term_reward, term, sizeof term);
    memcpy(t.step_reward, step_reward, sizeof step_reward);
    for (int i = 0; i < FT_NUM_CUM; i++) for (int k = 0; k < 4; k++) t.cum_next[i][k] = cum_next[i][k];

    // ---- BattleAI input sequences (BattleAI.cs:192-342): F = forward, B = backward, N = none ----
    enum { N = 0, F = 1, B = 2 };
    std::vector<uint8_t> mp;
    auto rep = [&](std::vector<uint8_t> &v, int val, int n) { for (int i = 0; i < n; i++) v.push_back((uint8_t)val); };
    auto dash = [&](std::vector<uint8_t> &v) { v.push_back(F); v.push_back(N); v.push_back(F); };  // :330-342 (both dashes tap FORWARD)
    int id = 1;
    uint16_t move_off[8] = {0}, move_len[8] = {0}, att_off[8] = {0}, att_len[8] = {0};
    auto begin = [&](std::vector<uint8_t> &v, uint16_t *off) { off[id] = (uint16_t)v.size(); };
    auto end = [&](std::vector<uint8_t> &v, uint16_t *off, uint16_t *len) { len[id] = (uint16_t)(v.size() - off[id]); id++; };
    begin(mp, move_off); rep(mp, N, 30); end(mp, move_off, move_len);                                   // 1 AddNeutralMovement
    begin(mp, move_off); rep(mp, F, 40); rep(mp, B, 10); rep(mp, F, 30); rep(mp, B, 10); end(mp, move_off, move_len); // 2 FarApproach1
    begin(mp, move_off); dash(mp); rep(mp, B, 25); dash(mp); rep(mp, B, 25); end(mp, move_off, move_len);             // 3 FarApproach2
    begin(mp, move_off); rep(mp, F, 30); rep(mp, B, 10); rep(mp, F, 20); rep(mp, B, 10); end(mp, move_off, move_len); // 4 MidApproach1
    begin(mp, move_off); dash(mp); rep(mp, B, 30); end(mp, move_off, move_len);                         // 5 MidApproach2
    begin(mp, move_off); rep(mp, B, 60); end(mp, move_off, move_len);                                   // 6 FallBack1
    begin(mp, move_off); dash(mp); rep(mp, B, 60); end(mp, move_off, move_len);                         // 7 FallBack2
    // InputDefine bits per side: P1 forward = Right (2), backward = Left (1); P2 mirrored (BattleAI.cs:380-390)
    for (size_t i = 0; i < mp.size() && i < (size_t)kMovePatBytes; i++) {
        t.move_pat[0][i] = mp[i] == F ? 2 : mp[i] == B ? 1 : 0;
        t.move_pat[1][i] = mp[i] == F ? 1 : mp[i] == B ? 2 : 0;
    }
    std::vector<uint8_t> apv;
    const int A = 4;
    id = 1;
    begin(apv, att_off); rep(apv, 0, 30); end(apv, att_off, att_len);                                   // 1 AddNoAttack
    begin(apv, att_off); rep(apv, A, 1); rep(apv, 0, 18); end(apv, att_off, att_len);                   // 2 OneHitImmediate
    begin(apv, att_off); rep(apv, A, 1); rep(apv, 0, 3); rep(apv, A, 1); rep(apv, 0, 18); end(apv, att_off, att_len); // 3 TwoHitImmediate
    begin(apv, att_off); rep(apv, A, 60); rep(apv, 0, 1); end(apv, att_off, att_len);                   // 4 ImmediateSpecial
    begin(apv, att_off); rep(apv, A, 120); rep(apv, 0, 1); end(apv, att_off, att_len);                  // 5 DelaySpecial
    memcpy(t.att_pat, apv.data(), apv.size());
    for (int i = 0; i < 8; i++) {
        t.move_meta[i] = move_off[i] | (uint32_t)move_len[i] << 16;
        t.att_meta[i] = att_off[i] | (uint32_t)att_len[i] << 16;
    }

    // ---- SelectMovement / SelectAttack per distance bucket (BattleAI.cs:68-190): Random.Range(0, n) picks the r-th
    //      nibble of `sel` = pattern id.  Buckets: 0 dist > 4, 1 > 3, 2 > 2.5, 3 > 2, 4 else. ----
    static const uint32_t n_m[5] = { 2, 7, 5, 4, 3 };
    static const uint32_t sel_m[5] = { 0x32u, 0x1325544u, 0x17654u, 0x1176u, 0x176u };
    static const uint32_t n_a[5] = { 4, 5, 3, 6, 3 };
    static const uint32_t sel_a[5] = { 0x1111u, 0x52211u, 0x321u, 0x543322u, 0x332u };
    for (int b = 0; b < 5; b++) {
        t.bot[b].n_m = n_m[b]; t.bot[b].sel_m = sel_m[b];
        t.bot[b].n_a = n_a[b] | (b == 1 ? 256u : 0u);   // bit 8: an opponent in a normal attack is answered with TwoHitImmediate (:150-160)
        t.bot[b].sel_a = sel_a[b];
        t.bot[b].magic_m = (~0ull) / (unsigned long long)n_m[b] + 1ull;   // floor(2^64 / n) + 1
        t.bot[b].magic_a = (~0ull) / (unsigned long long)n_a[b] + 1ull;
    }
    static const uint16_t dash_fsm[FT_NUM_DASH_STATES][4] = FT_DASH_FSM_INIT;
    static const uint8_t arun_lut[64 * 8] = FT_ARUN_LUT_INIT;
    static const uint8_t req_lut[2][1024] = FT_REQ_LUT_INIT;
    memcpy(t.req_lut, req_lut, sizeof req_lut);
    for (int i = 0; i < FT_NUM_DASH_STATES; i++) for (int d = 0; d < 4; d++) t.dash_fsm[i][d] = dash_fsm[i][d];
    memcpy(t.arun_lut, arun_lut, sizeof arun_lut);
    // index = clamp(ceil(2 * distance), 4, 9) - 4
    static const uint8_t bucket_of[6] = { 4, 3, 2, 1, 1, 0 };
    for (int i = 0; i < 6; i++) t.bucket_of[i] = bucket_of[i];
}

}  // namespace fg
#endif
