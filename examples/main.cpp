// main.cpp -- the reference's driver program (main.mojo:11-45) over the C ABI, in a compiled language.
//
// The reference's own language (Mojo 0.25.7) has no toolchain in this image, so this file stands where the
// Mojo shim would: it does exactly what main.mojo does -- Whisper(), WeightLoader(path) + load, Tokenizer(path),
// Tensor(80, 3000) filled from sample_input.bin, transcribe, print ids and text -- but through
// include/whisper_b200.h only (no Python, no torch: the library links against libcudart alone).
//
//   g++ -O2 -std=c++17 -Iinclude examples/main.cpp -Lwhisper_mojo_b200 -lwhisper_b200 -Wl,-rpath,... -o examples/main
//   examples/main [weights.bin [sample_input.bin [vocab.txt]]]      (defaults = main.mojo's file names)
//
// Extra (not in main.mojo): `--chunks N` transcribes N copies of the input as one batch, `--pcm` reads
// 480000 fp32 samples of 16 kHz audio instead of a log-mel and runs the device frontend first, and `--low-latency`
// selects the small-batch decode (option small_batch = 8) for one-clip-at-a-time use.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "whisper_b200.h"

static void die(const char *what) {
    char msg[512] = {0};
    wb_last_error(msg, sizeof msg);
    std::fprintf(stderr, "error: %s: %s\n", what, msg);
    std::exit(1);
}

// tokenizer.mojo:7-13: the vocabulary is the file split on '\n'
static std::vector<std::string> load_vocab(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    std::vector<std::string> v;
    if (!f) return v;
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    size_t a = 0;
    for (;;) {
        size_t b = s.find('\n', a);
        v.push_back(s.substr(a, b == std::string::npos ? std::string::npos : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return v;
}

static void replace_all(std::string &s, const std::string &from, const std::string &to) {
    for (size_t p = 0; (p = s.find(from, p)) != std::string::npos; p += to.size()) s.replace(p, from.size(), to);
}

// tokenizer.mojo:15-28
static std::string decode(const std::vector<std::string> &vocab, const int32_t *ids, int n) {
    std::string out;
    for (int i = 0; i < n; i++) {
        if (ids[i] < 0 || (size_t)ids[i] >= vocab.size()) continue;
        std::string t = vocab[ids[i]];
        const bool special = t.size() >= 4 && t.compare(0, 2, "<|") == 0 && t.compare(t.size() - 2, 2, "|>") == 0;
        if (special) continue;
        replace_all(t, "\xC4\xA0", " ");  // U+0120 'Ġ'
        replace_all(t, "\\n", "\n");
        out += t;
    }
    return out;
}

int main(int argc, char **argv) {
    std::string weights = "whisper_tiny_weights.bin", input = "sample_input.bin", vocab_path = "vocab.txt";
    int chunks = 1, pos = 0;
    bool pcm = false, low_latency = false;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--chunks") && i + 1 < argc) chunks = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--pcm")) pcm = true;
        else if (!std::strcmp(argv[i], "--low-latency")) low_latency = true;
        else if (pos == 0) weights = argv[i], pos++;
        else if (pos == 1) input = argv[i], pos++;
        else vocab_path = argv[i], pos++;
    }
    if (chunks < 1) chunks = 1;

    std::printf("Initializing Whisper Tiny on the B200 path...\n");
    wm_config cfg = {384, 6, 4, 51865, 1500, 448, 80, {50258, 50259, 50359, 50363}, 50257, 195, 1, {0, 0}};
    wm_model model = 0;
    if (wm_create(&cfg, nullptr, &model) != 0) die("wm_create");
    if (low_latency && wm_set_option(model, "small_batch", 8) != 0) die("wm_set_option");

    std::printf("Loading weights from %s...\n", weights.c_str());
    if (wm_load_weights_file(model, weights.c_str()) != 0) die("wm_load_weights_file");

    std::printf("Loading vocabulary from %s...\n", vocab_path.c_str());
    const std::vector<std::string> vocab = load_vocab(vocab_path);

    std::printf("Loading sample input from %s...\n", input.c_str());
    const size_t per = pcm ? (size_t)480000 : (size_t)80 * 3000;
    std::vector<float> in(per * chunks);
    {
        std::ifstream f(input, std::ios::binary);
        if (!f || !f.read(reinterpret_cast<char *>(in.data()), per * sizeof(float))) {
            std::fprintf(stderr, "error: %s does not hold %zu fp32 values\n", input.c_str(), per);
            return 1;
        }
        for (int c = 1; c < chunks; c++) std::memcpy(in.data() + c * per, in.data(), per * sizeof(float));
    }

    const int T = 5 + cfg.max_iters;  // 4 prompt ids + 1 + 195 (whisper.mojo:200-221)
    std::vector<int32_t> toks((size_t)chunks * T), lens(chunks);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = pcm ? wm_transcribe_pcm(model, in.data(), chunks, toks.data(), lens.data())
                       : wm_transcribe(model, in.data(), chunks, toks.data(), lens.data());
    const auto t1 = std::chrono::steady_clock::now();
    if (rc != 0) die("wm_transcribe");

    std::printf("Transcription time:  %.6f seconds\n", std::chrono::duration<double>(t1 - t0).count());
    std::printf("\nToken IDs:\n");
    for (int i = 0; i < lens[0]; i++) std::printf("%d ", toks[i]);
    std::printf("\n");
    for (int c = 1; c < chunks; c++)
        if (lens[c] != lens[0] || std::memcmp(&toks[(size_t)c * T], toks.data(), lens[0] * sizeof(int32_t)) != 0) {
            std::fprintf(stderr, "error: chunk %d of the replicated batch differs from chunk 0\n", c);
            return 2;
        }
    if (!vocab.empty()) {
        const std::string text = decode(vocab, toks.data(), lens[0]);
        std::printf("\n========================================\nFINAL TRANSCRIPTION:\n");
        std::printf("========================================\n%s\n========================================\n", text.c_str());
    }
    std::printf("\nDone.\n");
    wm_destroy(model);
    return 0;
}
