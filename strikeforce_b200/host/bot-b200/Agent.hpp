// bots/bot-b200/Agent.hpp -- the reference's bot plugin surface on a BATCH of arenas.
//
// The reference selects a bot folder at compile time (StrikeForce-client/selected_agent.hpp:25,
// selected_custom.hpp:25, README.md:288-300); the folder provides `class Agent` with
//     int  predict(const std::vector<float>& obs);     (bots/bot-0.5/Agent.hpp:178)
//     void update(int action, bool imitate);            (:217)
//     bool in_training();  bool is_manual();            (:236-268; bots/bot-0/Agent.hpp:27-37)
// and the out-of-class gameplay::prepare / bot / view (Custom.hpp).  This is the same surface with
// one row per (arena, driven human): observations arrive as an fp32 DEVICE tensor [B, 32, 31, 31]
// written by sf_observe, predict returns int64 [B] indices into gameplay::action, and nothing
// round-trips through the host.  What maps observations to action probabilities is a `Policy`
// the host plugs in (the reference hard-wires its AgentModel; a libtorch module, a TorchScript
// file or anything else that eats a device tensor fits).  Header-only; needs libtorch.
#pragma once
#include <ATen/cuda/CUDAGeneratorImpl.h>
#include <torch/torch.h>

#include <memory>
#include <stdexcept>

namespace sfb200 {

// observations [B, 32, 31, 31] (device, fp32) -> action probabilities [B, n_actions] (device, fp32)
struct Policy {
    virtual ~Policy() = default;
    virtual torch::Tensor probabilities(const torch::Tensor &obs) = 0;
    // the actions just taken (one-hot feedback of Agent::update, bots/bot-0.5/Agent.hpp:221-223)
    virtual void chosen(const torch::Tensor & /*actions*/) {}
    // a new game for the rows of `mask` (the reference builds a new Agent per game: reset_memory)
    virtual void new_games(const torch::Tensor & /*mask*/) {}
};

// bots/bot-0/Agent.hpp:27-37, the do-nothing template: always action 0 ('+')
struct IdlePolicy : Policy {
    int n_actions;
    explicit IdlePolicy(int n = 9) : n_actions(n) {}
    torch::Tensor probabilities(const torch::Tensor &obs) override
    {
        auto p = torch::zeros({obs.size(0), n_actions}, obs.options());
        p.select(1, 0).fill_(1.0f);
        return p;
    }
};

class Agent {
public:
    // greedy = false: sample from the distribution like the reference (std::discrete_distribution,
    // bots/bot-0.5/Agent.hpp:210-212; here with a seeded device generator); true: argmax
    explicit Agent(std::shared_ptr<Policy> policy, bool training = true, bool greedy = false, uint64_t seed = 0)
        : policy_(std::move(policy)), training_(training), greedy_(greedy), seed_(seed)
    {
        if (!policy_) throw std::invalid_argument("Agent: no policy");
    }
    // Agent::predict for every row: int64 [B] on the device of obs
    torch::Tensor predict(const torch::Tensor &obs)
    {
        torch::NoGradGuard ng;
        last_p_ = policy_->probabilities(obs);
        if (greedy_) return last_p_.argmax(1);
        if (!gen_.defined()) {
            gen_ = at::cuda::detail::createCUDAGenerator(obs.device().index());
            gen_.set_current_seed(seed_);
        }
        return torch::multinomial(last_p_, 1, false, gen_).view({-1});
    }
    void update(const torch::Tensor &actions, bool /*imitate*/) { policy_->chosen(actions); }
    void new_games(const torch::Tensor &mask) { policy_->new_games(mask); }
    bool in_training() const { return training_; }
    bool is_manual() const { return false; }
    const torch::Tensor &last_probabilities() const { return last_p_; }

private:
    std::shared_ptr<Policy> policy_;
    bool training_, greedy_;
    uint64_t seed_;
    at::Generator gen_;
    torch::Tensor last_p_;
};

} // namespace sfb200
