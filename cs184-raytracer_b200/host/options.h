// options.h — command-line surface of `as2`, identical to the reference's
// (src/options.h:5-30, src/options.cpp:7-90): -t/--threads, -w/--width, -h/--height
// (-h is HEIGHT, not help), -o/--output, --bdepth, --intersection-only, --help,
// positional .rti files; one global instance `programOptions` that the render path
// reads, exactly as src/scene.cpp:31,39,50,69 do.  Extra (ours): --brute-force, and --aa N
// (N x N supersampling, the reference's stated next feature, TODO:2), --gpus N (render on N GPUs of this
// box: rt_params.n_gpus; default 1, or the AS2_GPUS environment variable).
#pragma once
#include <string>
#include <vector>

namespace as2 {

class Options {
public:
    bool parseCommandLine(int argc, char* argv[]);
    void printHelp(const char* prog);

    std::vector<std::string> inputFilenames_;
    std::string outputFilename_;
    int renderThreadsCount_ = 1;   // accepted for compatibility; the GPU path ignores it
    int renderWidth_ = 500;
    int renderHeight_ = 500;
    int bounceDepth_ = 10;
    bool intersectionOnly_ = false;
    bool bruteForce_ = false;      // --brute-force: debug aid, skips the LBVH
    int samples_ = 1;              // --aa N: N x N rays per pixel, averaged (1 = the reference's single centre ray)
    int gpus_ = 1;                 // --gpus N / AS2_GPUS: GPUs of this box that share the frame's tiles
};

extern Options programOptions;

}  // namespace as2
