// Golden-vector generator for SURVEY.md 8f rank 4 (the gradient hub): what
// AgentServer::aggregate_and_update + compute_update_vector
// (reference bots/bot-0.5/AgentServer.cpp:465-524) do to the server's model when clients send
// gradients -- sum over the clients in order, divide by their number, one torch::optim::AdamW step
// with AdamWOptions(lr) (:101-105), update vector = theta_new - theta_old -- restated here around
// the reference's OWN AgentModel (bots/bot-0.5/Modules.hpp, unmodified) and libtorch's own AdamW.
// (The server itself cannot be driven headless: it is a TCP accept loop around exactly these lines.)
// Parameters and gradients come from a closed integer-hash formula, so the fixture only holds the
// update vectors.  Test infrastructure only; built by build_hub_oracle.sh into oracle/_ref/.
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

static float hash_unit(uint32_t i, uint32_t k)
{
    uint32_t u = i * 2654435761u + k * 40503u + 12345u;
    u ^= u >> 15;
    u *= 2246822519u;
    u ^= u >> 13;
    return (float)(u >> 8) / 16777216.0f;
}

static torch::Tensor formula(const torch::Tensor &like, uint32_t k, float scale)
{
    torch::Tensor w = torch::empty({like.numel()});
    float *d = w.data_ptr<float>();
    for (int64_t i = 0; i < w.numel(); ++i) d[i] = (hash_unit((uint32_t)i, k) - 0.5f) * scale;
    return w.view(like.sizes());
}

int main(int argc, char **argv)
{
    const char *out = argc > 1 ? argv[1] : "hub_golden.bin";
    const int hidden = argc > 2 ? atoi(argv[2]) : 8, rounds = argc > 3 ? atoi(argv[3]) : 3, clients = argc > 4 ? atoi(argv[4]) : 2;
    const double lr = 1e-3; // AgentServer.cpp:614
    AgentModel model(32, 31, 31, hidden, 9);
    {
        torch::NoGradGuard ng;
        uint32_t k = 0;
        for (auto &p : model->parameters()) p.copy_(formula(p, k++, 0.16f));
    }
    torch::optim::AdamW optimizer(model->parameters(), torch::optim::AdamWOptions(lr)); // :102-105
    std::ofstream f(out, std::ios::binary);
    const int sz = (int)model->parameters().size();
    for (int r = 0; r < rounds; ++r) {
        // what the clients sent (send_gradient, AgentClient.hpp:72-82)
        std::vector<std::vector<torch::Tensor>> grads(clients);
        for (int c = 0; c < clients; ++c)
            for (int i = 0; i < sz; ++i)
                grads[c].push_back(formula(model->parameters()[i], 5000u + 1000u * (uint32_t)r + 100u * (uint32_t)c + (uint32_t)i, 0.02f));
        // aggregate_and_update(), :480-511
        std::vector<torch::Tensor> avg, theta_old, theta_new;
        for (int i = 0; i < sz; ++i) {
            torch::Tensor sum_grad = torch::zeros_like(model->parameters()[i]);
            for (int c = 0; c < clients; ++c) sum_grad += grads[c][i];
            avg.push_back(sum_grad / clients);
        }
        for (int i = 0; i < sz; ++i) theta_old.push_back(model->parameters()[i].clone().detach());
        optimizer.zero_grad();
        for (int i = 0; i < sz; ++i) model->parameters()[i].mutable_grad() = avg[i].clone().detach();
        optimizer.step();
        for (int i = 0; i < sz; ++i) theta_new.push_back(model->parameters()[i].clone().detach());
        // compute_update_vector(), :513-524
        for (int i = 0; i < sz; ++i) put(f, "u:" + std::to_string(r) + ":" + std::to_string(i), (theta_new[i] - theta_old[i]).detach().clone());
    }
    f.close();
    std::printf("wrote %s (hidden %d, %d rounds, %d clients, %d tensors)\n", out, hidden, rounds, clients, sz);
    return 0;
}
