// b200_play -- the loop of gameplay::play() for a batch of arenas, hosted in C++ above the C ABI
// (include/strikeforce_b200.h) through the bot plugin surface of bot-b200/{Agent,Custom}.hpp.
//
//   b200_play <config blob> <ticks> [idle|uniform]
//
// <config blob> is written by strikeforce_b200.config.dump_config (or by any host that fills an
// sf_config).  "idle" is the reference's template agent (bots/bot-0/Agent.hpp: always '+'), "uniform"
// a uniformly random policy over "+xzqeawsd".  Prints one JSON line with the device-reduced episode
// statistics.  There is no CPU path: without a CUDA device sf_create fails with SF_ERR_NO_DEVICE.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bot-b200/Custom.hpp"

namespace {
struct UniformPolicy : sfb200::Policy {
    torch::Tensor probabilities(const torch::Tensor &obs) override
    {
        return torch::full({obs.size(0), 9}, 1.0f / 9.0f, obs.options());
    }
};
} // namespace

int main(int argc, char **argv)
{
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <config blob> <ticks> [idle|uniform]\n", argv[0]);
        return 2;
    }
    try {
        sfb200::Config c = sfb200::Config::load(argv[1]);
        const int ticks = std::atoi(argv[2]);
        const bool uniform = argc > 3 && !std::strcmp(argv[3], "uniform");
        std::shared_ptr<sfb200::Policy> policy;
        if (uniform) policy = std::make_shared<UniformPolicy>();
        else policy = std::make_shared<sfb200::IdlePolicy>();
        auto agent = std::make_shared<sfb200::Agent>(policy, /*training=*/false, /*greedy=*/!uniform, /*seed=*/1);
        sfb200::BatchedGameplay g(c.cfg, agent);
        for (int t = 0; t < ticks; ++t) g.tick();
        auto st = g.stats();
        unsigned long long x = 0;
        for (uint64_t h : g.state_hash()) x ^= h;
        std::printf("{\"arenas\": %d, \"ticks\": %d, \"steps\": %lld, \"episodes\": %lld, \"wins\": %lld, \"deaths\": %lld, "
                    "\"kills\": %lld, \"rng_draws\": %lld, \"launches\": %lld, \"hash_xor\": %llu}\n",
                    g.n_envs(), ticks, (long long)st[SF_STAT_STEPS], (long long)st[SF_STAT_EPISODES], (long long)st[SF_STAT_WINS],
                    (long long)st[SF_STAT_DEATHS], (long long)st[SF_STAT_KILLS], (long long)st[SF_STAT_RNG_DRAWS],
                    (long long)sf_launch_count(g.handle()), x);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "b200_play: %s\n", e.what());
        return 1;
    }
    return 0;
}
