// bots/bot-b200/Custom.hpp -- gameplay::prepare / bot / view (bots/bot-0.5/Custom.hpp:137-168) and the
// loop of gameplay::play() (gameplay.hpp:1443-1472) for every arena of an sf_handle.
//
// The engine behind it is the C ABI of include/strikeforce_b200.h (libstrikeforce_b200.so): this file
// is the host side of that boundary in the reference's own language.  Observations, predictions and
// commands stay on the device; all work goes to the current CUDA stream of the calling thread.
#pragma once
#include <c10/cuda/CUDAStream.h>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Agent.hpp"
extern "C" {
#include "strikeforce_b200.h"
}

namespace sfb200 {

// An sf_config with the arrays it points to.  load() reads the blob that
// strikeforce_b200.config.dump_config() writes: the struct, then map_cells[SF_CELLS], then
// map_portal[SF_CELLS] (a host that parses map/, Items/ and character/ itself fills the struct directly).
struct Config {
    sf_config cfg{};
    std::vector<uint8_t> cells;
    std::vector<int16_t> portal;
    static Config load(const std::string &path)
    {
        Config c;
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("cannot open " + path);
        c.cells.resize(SF_CELLS), c.portal.resize(SF_CELLS);
        f.read(reinterpret_cast<char *>(&c.cfg), sizeof c.cfg);
        f.read(reinterpret_cast<char *>(c.cells.data()), SF_CELLS);
        f.read(reinterpret_cast<char *>(c.portal.data()), SF_CELLS * 2);
        if (!f) throw std::runtime_error(path + ": truncated configuration blob");
        if (c.cfg.abi_version != SF_ABI_VERSION) throw std::runtime_error(path + ": written for another ABI version");
        c.cfg.map_cells = c.cells.data(), c.cfg.map_portal = c.portal.data();
        return c;
    }
};

class BatchedGameplay {
public:
    // gameplay::prepare(Human&): the action table and the agent (Custom.hpp:161-165)
    // channels_last: the observations are written channel-innermost (SF_OBS_NHWC) and handed to the agent as a
    // [B, 32, 31, 31] tensor with channels-last strides -- the same values, and a convolution reads them
    // without transposing 123,008 bytes per observation first
    BatchedGameplay(const sf_config &cfg, std::shared_ptr<Agent> agent, const std::string &action = "+xzqeawsd",
                    bool channels_last = false)
        : agent_(std::move(agent)), channels_last_(channels_last)
    {
        if (sf_create(&cfg, &h_) != SF_OK) throw std::runtime_error(std::string("sf_create: ") + sf_last_error(nullptr));
        n_ = sf_num_envs(h_), a_ = sf_agents_per_env(h_);
        auto dev = torch::Device(torch::kCUDA, c10::cuda::current_device());
        actions_ = torch::full({n_, a_}, (int)'+', torch::dtype(torch::kUInt8).device(dev));
        table_ = torch::tensor(std::vector<uint8_t>(action.begin(), action.end()), torch::dtype(torch::kUInt8)).to(dev);
        step_out_ = torch::empty({n_, 8}, torch::dtype(torch::kInt32).device(dev));
    }
    ~BatchedGameplay()
    {
        if (h_) sf_destroy(h_);
    }
    BatchedGameplay(const BatchedGameplay &) = delete;
    BatchedGameplay &operator=(const BatchedGameplay &) = delete;

    int n_envs() const { return n_; }
    int n_agents() const { return a_; }
    sf_handle *handle() { return h_; }
    torch::Tensor &actions() { return actions_; }
    const torch::Tensor &observations() const { return obs_; }

    // gameplay::bot(Human&) for the humans selected in agent_mask, of EVERY arena: observe -> predict ->
    // table look-up; returns uint8 [n_envs, n_selected] command symbols (get_my_action, gameplay.hpp:956)
    torch::Tensor bot(uint32_t agent_mask = 1u, int phase = SF_OBS_P1)
    {
        const int nsel = __builtin_popcount(agent_mask);
        if (!obs_.defined() || obs_.size(0) != (int64_t)n_ * nsel) {
            auto opt = torch::dtype(torch::kFloat32).device(actions_.device());
            obs_ = channels_last_ ? torch::empty({(int64_t)n_ * nsel, SF_OBS_WIN, SF_OBS_WIN, SF_OBS_CH}, opt).permute({0, 3, 1, 2})
                                  : torch::empty({(int64_t)n_ * nsel, SF_OBS_CH, SF_OBS_WIN, SF_OBS_WIN}, opt);
        }
        check(sf_observe(h_, obs_.data_ptr<float>(), phase | (channels_last_ ? SF_OBS_NHWC : 0), agent_mask, stream()), "sf_observe");
        torch::Tensor idx = agent_->predict(obs_);
        agent_->update(idx, false);
        return table_.index_select(0, idx).view({n_, nsel});
    }
    void view() {} // gameplay::view(): a no-op in every shipped bot (bots/bot-0.5/Custom.hpp:167)

    // one iteration of the loop of gameplay::play() for all arenas: the player's command from the agent
    // (P1); with agent-driven squad humans the step is split so that they observe at P2 (gameplay.hpp:933)
    void tick()
    {
        const uint32_t squad = ((1u << a_) - 1u) & ~1u;
        actions_.select(1, 0).copy_(bot(1u, SF_OBS_P1).select(1, 0));
        if (squad) {
            check(sf_step_a(h_, stream()), "sf_step_a");
            actions_.slice(1, 1, a_).copy_(bot(squad, SF_OBS_P2));
            check(sf_step_b(h_, actions_.data_ptr<uint8_t>(), stream()), "sf_step_b");
        } else {
            check(sf_step(h_, actions_.data_ptr<uint8_t>(), stream()), "sf_step");
        }
        view();
        check(sf_get(h_, SF_FIELD_STEP_OUT, step_out_.data_ptr<int32_t>(), stream()), "sf_get");
        agent_->new_games(step_out_.select(1, 0).ne(SF_RUNNING).repeat_interleave(agent_rows_per_env_));
    }
    // rows the agent keeps per arena (1: the player only)
    void set_agent_rows_per_env(int r) { agent_rows_per_env_ = r; }

    std::vector<int64_t> stats()
    {
        auto t = torch::empty({SF_STAT_COUNT}, torch::dtype(torch::kInt64).device(actions_.device()));
        check(sf_get(h_, SF_FIELD_STATS, t.data_ptr<int64_t>(), stream()), "sf_get");
        auto c = t.cpu();
        return std::vector<int64_t>(c.data_ptr<int64_t>(), c.data_ptr<int64_t>() + SF_STAT_COUNT);
    }
    std::vector<uint64_t> state_hash()
    {
        auto t = torch::empty({n_}, torch::dtype(torch::kInt64).device(actions_.device()));
        check(sf_get(h_, SF_FIELD_STATE_HASH, t.data_ptr<int64_t>(), stream()), "sf_get");
        auto c = t.cpu();
        const uint64_t *p = reinterpret_cast<const uint64_t *>(c.data_ptr<int64_t>());
        return std::vector<uint64_t>(p, p + n_);
    }

private:
    static void *stream() { return c10::cuda::getCurrentCUDAStream().stream(); }
    void check(int rc, const char *what)
    {
        if (rc != SF_OK) throw std::runtime_error(std::string(what) + ": " + sf_last_error(h_));
    }
    sf_handle *h_ = nullptr;
    std::shared_ptr<Agent> agent_;
    int n_ = 0, a_ = 0, agent_rows_per_env_ = 1;
    torch::Tensor obs_, actions_, table_, step_out_;
    bool channels_last_ = false;
};

} // namespace sfb200
