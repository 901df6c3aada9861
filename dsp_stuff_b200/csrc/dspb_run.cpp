// dspb_run — headless driver (the reference's main.rs/runtime.rs minus GUI and cpal): loads a graph saved by
// dsp-stuff (DSPConfig JSON, runtime.rs:44-48), instantiates it over C channels on one GPU, streams synthetic
// noise through it in device blocks and prints throughput plus a checksum of the last output block.
//   dspb_run graph.json [channels=1024] [block=1024] [blocks_per_call=16] [seconds=2.0] [device=0]
// Links against libdspb200.so only (the C ABI in include/dspb200.h); host buffers, DSPB_MEM_HOST.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/dspb200.h"

static uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// SURVEY.md section 8d noise: uniform in [-0.5, 0.5), exactly representable in f32
static float noise(uint64_t seed, uint64_t c, uint64_t n) {
    const uint64_t h = splitmix64(seed ^ ((c << 32) | n));
    return (float)((int64_t)(h >> 40) - (1 << 23)) * (1.0f / 8388608.0f) * 0.5f;
}

#define CK(call)                                                                  \
    do {                                                                          \
        int rc_ = (call);                                                         \
        if (rc_ != DSPB_OK) {                                                     \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, dspb_last_error()); \
            return 1;                                                             \
        }                                                                         \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s graph.json [channels] [block] [blocks_per_call] [seconds] [device]\n", argv[0]);
        return 2;
    }
    const int C = argc > 2 ? atoi(argv[2]) : 1024;
    const int block = argc > 3 ? atoi(argv[3]) : 1024;
    const int bpc = argc > 4 ? atoi(argv[4]) : 16;
    const double seconds = argc > 5 ? atof(argv[5]) : 2.0;
    const int device = argc > 6 ? atoi(argv[6]) : 0;
    std::ifstream f(argv[1]);
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string json = ss.str();

    dspb_config cfg{};
    cfg.channels = C;
    cfg.block = block;
    cfg.device = device;
    cfg.max_samples = (int64_t)block * bpc;
    dspb_engine* e = nullptr;
    CK(dspb_engine_create(&cfg, &e));
    CK(dspb_load_graph_json(e, json.c_str()));
    int64_t n_in = 0, n_out = 0;
    {   // count terminals from the JSON text (typename strings), same order as the engine
        for (size_t p = 0; (p = json.find("\"typename\"", p)) != std::string::npos; p++) {
            const size_t q = json.find('"', json.find(':', p) + 1);
            const std::string t = json.substr(q + 1, json.find('"', q + 1) - q - 1);
            if (t == "input") n_in++;
            if (t == "output") n_out++;
        }
    }
    std::vector<char> plan((size_t)dspb_describe_plan(e, nullptr, 0));
    dspb_describe_plan(e, plan.data(), (int64_t)plan.size());
    fprintf(stderr, "%s", plan.data());

    const int64_t n = (int64_t)block * bpc;
    std::vector<std::vector<float>> in(n_in, std::vector<float>((size_t)C * n)), out(n_out, std::vector<float>((size_t)C * n));
    std::vector<const float*> ip;
    std::vector<float*> op;
    for (auto& v : in) ip.push_back(v.data());
    for (auto& v : out) op.push_back(v.data());
    const int64_t calls = (int64_t)(seconds * 48000.0 / (double)n) + 1;
    double busy = 0.0;
    for (int64_t k = 0; k < calls; k++) {
        for (int64_t t = 0; t < n_in; t++)
            for (int c = 0; c < C; c++)
                for (int64_t i = 0; i < n; i++) in[t][(size_t)c * n + i] = noise(42 + t, (uint64_t)c, (uint64_t)(k * n + i));
        const auto t0 = std::chrono::steady_clock::now();
        CK(dspb_process(e, ip.data(), op.data(), n, DSPB_MEM_HOST, nullptr));
        busy += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    double sum = 0.0;
    for (auto& v : out)
        for (float x : v) sum += (double)x;
    printf("{\"channels\": %d, \"samples_per_channel\": %lld, \"seconds_in_engine\": %.6f, \"channel_samples_per_sec\": %.6g, "
           "\"x_realtime_per_channel\": %.3f, \"checksum_last_block\": %.9g}\n",
           C, (long long)(calls * n), busy, (double)C * (double)(calls * n) / busy, (double)(calls * n) / 48000.0 / busy, sum);
    dspb_engine_destroy(e);
    return 0;
}
