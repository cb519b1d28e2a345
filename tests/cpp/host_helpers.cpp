// CPU-only check of the C++ host mirror's file helpers (no ss_ctx is created): Audacity labels, the 32-bit WAV writer
// and the PCM reader. Prints one JSON line that tests/test_host_cpu.py compares with the Python mirror's results.
#include <cstdio>

#include "soundsym.hpp"

using namespace soundsym;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string labels = argv[1], wav = argv[2];
    auto ts = audacity_labels_to_timestamps(labels);
    printf("{\"n\": %zu, \"stamps\": [", ts.size());
    for (size_t i = 0; i < ts.size(); i++)
        printf("%s[%.17g, %.17g, %s%s%s]", i ? ", " : "", ts[i].start, ts[i].end, ts[i].label ? "\"" : "", ts[i].label ? ts[i].label->c_str() : "null",
               ts[i].label ? "\"" : "");
    const double s[] = {0.0, 0.5, -0.5, 1.0, -1.0, 2.0, -2.0, std::numeric_limits<double>::quiet_NaN(), 0.123456789};
    Sound::write_wav_i32(wav, s, sizeof(s) / sizeof(s[0]), 22050.0);
    double sr = 0;
    std::vector<double> back = Sound::read_wav_samples(wav, &sr);
    printf("], \"sr\": %.1f, \"back\": [", sr);
    for (size_t i = 0; i < back.size(); i++) printf("%s%.17g", i ? ", " : "", back[i]);
    printf("]}\n");
    return 0;
}
