// Golden-vector generator for SURVEY.md 8f rank 2: the reference's OWN policy network
// (bots/bot-0.5/Modules.hpp, AgentModel: CNN + 2 GRUs + residual heads), unmodified, run with
// libtorch on the CPU.  Writes every parameter, a sequence of observations and the (p, v) it
// returns for them (the GRU state is carried from call to call, as in Agent::predict).
// Test infrastructure only; built by oracle/ref_harness/build_policy_oracle.sh into oracle/_ref/.
#include "bots/bot-0.5/Modules.hpp"

#include <cstdio>
#include <fstream>

static void put(std::ofstream &f, const std::string &name, const torch::Tensor &t_)
{
    torch::Tensor t = t_.detach().contiguous().to(torch::kFloat32);
    int32_t nlen = (int32_t)name.size(), nd = (int32_t)t.dim();
    f.write((const char *)&nlen, 4);
    f.write(name.data(), nlen);
    f.write((const char *)&nd, 4);
    for (int i = 0; i < nd; ++i) {
        int64_t s = t.size(i);
        f.write((const char *)&s, 8);
    }
    f.write((const char *)t.data_ptr<float>(), (std::streamsize)(t.numel() * 4));
}

// Parameters and observations come from a closed formula (32-bit integer hash -> float32), so the
// fixture only has to hold the network's OUTPUTS: the test regenerates the same inputs with numpy.
static float hash_unit(uint32_t i, uint32_t k)
{
    uint32_t u = i * 2654435761u + k * 40503u + 12345u;
    u ^= u >> 15;
    u *= 2246822519u;
    u ^= u >> 13;
    return (float)(u >> 8) / 16777216.0f; // exact: 24 bits
}

int main(int argc, char **argv)
{
    const char *out = argc > 1 ? argv[1] : "policy_golden.bin";
    int steps = argc > 2 ? atoi(argv[2]) : 6;
    AgentModel model(32, 31, 31, 160, 9);
    model->eval();
    torch::NoGradGuard ng;
    std::ofstream f(out, std::ios::binary);
    uint32_t k = 0;
    for (auto &p : model->named_parameters()) {
        torch::Tensor w = torch::empty({p.value().numel()});
        float *d = w.data_ptr<float>();
        for (int64_t i = 0; i < w.numel(); ++i) d[i] = (hash_unit((uint32_t)i, k) - 0.5f) * 0.16f;
        p.value().copy_(w.view(p.value().sizes()));
        ++k;
    }
    for (int s = 0; s < steps; ++s) {
        // sparse non-negative inputs like real observations (8% of the entries in [0, 1.5))
        torch::Tensor x = torch::zeros({32 * 31 * 31});
        float *d = x.data_ptr<float>();
        for (int64_t i = 0; i < x.numel(); ++i)
            if (hash_unit((uint32_t)i, 1000u + (uint32_t)s) < 0.08f) d[i] = hash_unit((uint32_t)i, 2000u + (uint32_t)s) * 1.5f;
        x = x.view({1, 32, 31, 31});
        auto r = model->forward(x);
        put(f, "p:" + std::to_string(s), r[0]);
        put(f, "v:" + std::to_string(s), r[1]);
        // Agent::predict feeds the chosen action back as a one-hot (bots/bot-0.5/Agent.hpp:216-219)
        torch::Tensor one_hot = torch::zeros({9});
        one_hot[(s * 4 + 1) % 9] += 1;
        model->update_actions(one_hot);
    }
    f.close();
    std::printf("wrote %s (%d steps)\n", out, steps);
    return 0;
}
