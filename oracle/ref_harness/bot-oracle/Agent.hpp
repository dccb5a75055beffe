// Oracle bot: the `class Agent` that the reference's plugin surface expects
// (reference bots/bot-0/Agent.hpp:27-37 is the minimal template; the full contract is
// bots/bot-0.5/Agent.hpp:36-377).  Test infrastructure only.
//
// predict() does not run a network: it records the observation the reference's own
// gameplay::bot() (bots/bot-0.5/Custom.hpp:137-159) built, and returns the action index
// the harness queued for the human this agent is attached to.
#pragma once
#include "../../basic.hpp"
#include <cmath>

struct OracleAgentHub {
    static constexpr int OBS_LEN = 32 * 31 * 31;
    int next_action[64];          // per human slot: index into gameplay::action
    int capture[64];              // per human slot: capture the observation?
    int captured[64];             // per human slot: how many observations captured this step
    float obs[64][OBS_LEN];       // last captured observation per slot
    OracleAgentHub() { std::memset(this, 0, sizeof(*this)); }
};
inline OracleAgentHub g_hub;

class Agent {
public:
    int slot = 0;                 // human slot this agent drives (set by the harness)
    Agent(bool = true) {}
    int predict(const std::vector<float> &o) {
        if (slot >= 0 && slot < 64) {
            if (g_hub.capture[slot] && (int)o.size() == OracleAgentHub::OBS_LEN) {
                std::memcpy(g_hub.obs[slot], o.data(), sizeof(float) * OracleAgentHub::OBS_LEN);
                ++g_hub.captured[slot];
            }
            return g_hub.next_action[slot];
        }
        return 0;
    }
    void update(int, bool) {}
    bool in_training() { return false; }
    bool is_manual() { return false; }
};
