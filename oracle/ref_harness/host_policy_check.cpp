// The C++ host binding (strikeforce_b200/host/bot-b200/{Agent,Custom}.hpp over the C ABI) driven by the
// reference's OWN policy network: one unmodified AgentModel (bots/bot-0.5/Modules.hpp) per arena,
// evaluated on the DEVICE tensors sf_observe fills, greedy actions fed back into sf_step.
// Writes, per tick, the probabilities, the chosen command symbols and the canonical-state hash of
// every arena; tests/test_cpp_host.py replays the same arenas through the Python mirror
// (strikeforce_b200.bots / policy) and compares.  TEST INFRASTRUCTURE: includes the reference's
// header, so it is built into oracle/_ref/ by build_host_check.sh and never shipped.
//
//   host_policy_check <config blob> <ticks> <out file> [nhwc]     (nhwc: observations written channel-innermost)
#include "bots/bot-0.5/Modules.hpp"

#include <cstdio>
#include <fstream>

#include "bot-b200/Custom.hpp"

static float hash_unit(uint32_t i, uint32_t k)
{
    uint32_t u = i * 2654435761u + k * 40503u + 12345u;
    u ^= u >> 15;
    u *= 2246822519u;
    u ^= u >> 13;
    return (float)(u >> 8) / 16777216.0f;
}

// one reference network per row: the reference keeps the GRU state and the last action inside the
// module (reset_memory / update_actions, Modules.hpp:94-103), one Agent per game
struct ReferencePolicy : sfb200::Policy {
    std::vector<AgentModel> models;
    ReferencePolicy(int rows, torch::Device dev)
    {
        for (int r = 0; r < rows; ++r) {
            AgentModel m(32, 31, 31, 160, 9);
            torch::NoGradGuard ng;
            uint32_t k = 0;
            for (auto &p : m->named_parameters()) {
                torch::Tensor w = torch::empty({p.value().numel()});
                float *d = w.data_ptr<float>();
                for (int64_t i = 0; i < w.numel(); ++i) d[i] = (hash_unit((uint32_t)i, k) - 0.5f) * 0.16f;
                p.value().copy_(w.view(p.value().sizes()));
                ++k;
            }
            m->to(dev);
            m->eval();
            m->reset_memory();
            // reset_memory() builds its state on the CPU (the reference never leaves it); same values, on the device
            m->backbone->action_input = m->backbone->action_input.to(dev);
            m->backbone->h_state[0] = m->backbone->h_state[0].to(dev);
            m->backbone->h_state[1] = m->backbone->h_state[1].to(dev);
            models.push_back(m);
        }
    }
    torch::Tensor probabilities(const torch::Tensor &obs) override
    {
        std::vector<torch::Tensor> rows;
        for (size_t r = 0; r < models.size(); ++r) rows.push_back(models[r]->forward(obs.slice(0, r, r + 1))[0].view({1, 9}));
        return torch::cat(rows, 0);
    }
    void chosen(const torch::Tensor &actions) override
    {
        auto a = actions.cpu();
        for (size_t r = 0; r < models.size(); ++r) {
            torch::Tensor one_hot = torch::zeros({9}, torch::device(models[r]->parameters()[0].device()));
            one_hot[a[r].item<int64_t>()] += 1;
            models[r]->update_actions(one_hot); // Agent::update, bots/bot-0.5/Agent.hpp:221-223
        }
    }
};

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    at::globalContext().setAllowTF32CuBLAS(false);
    at::globalContext().setAllowTF32CuDNN(false);
    try {
        sfb200::Config c = sfb200::Config::load(argv[1]);
        const int ticks = atoi(argv[2]);
        auto dev = torch::Device(torch::kCUDA, 0);
        auto policy = std::make_shared<ReferencePolicy>(c.cfg.n_envs, dev);
        auto agent = std::make_shared<sfb200::Agent>(policy, false, /*greedy=*/true);
        const bool nhwc = argc > 4 && std::string(argv[4]) == "nhwc";
        sfb200::BatchedGameplay g(c.cfg, agent, "+xzqeawsd", nhwc);
        std::ofstream f(argv[3], std::ios::binary);
        for (int t = 0; t < ticks; ++t) {
            g.tick();
            auto p = agent->last_probabilities().cpu().contiguous();
            auto a = g.actions().cpu().contiguous();
            auto h = g.state_hash();
            f.write((const char *)p.data_ptr<float>(), p.numel() * 4);
            f.write((const char *)a.data_ptr<uint8_t>(), a.numel());
            f.write((const char *)h.data(), h.size() * 8);
        }
        std::printf("host_policy_check: %d arenas x %d ticks\n", g.n_envs(), ticks);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "host_policy_check: %s\n", e.what());
        return 1;
    }
    return 0;
}
