// options.cpp — see options.h.  Messages follow src/options.cpp:32-36,47-51,59-64,78-83,89.
#include "options.h"

#include <getopt.h>

#include <cerrno>
#include <climits>
#include <cstdlib>
#include <iostream>

namespace as2 {

Options programOptions;

namespace {
enum { kHelp = 1000, kBounceDepth, kIntersectionOnly, kBruteForce, kSamples, kGpus };
const struct option kLongOptions[] = {
    {"help", no_argument, nullptr, kHelp},
    {"output", required_argument, nullptr, 'o'},
    {"threads", required_argument, nullptr, 't'},
    {"width", required_argument, nullptr, 'w'},
    {"height", required_argument, nullptr, 'h'},
    {"bdepth", required_argument, nullptr, kBounceDepth},
    {"intersection-only", no_argument, nullptr, kIntersectionOnly},
    {"brute-force", no_argument, nullptr, kBruteForce},
    {"aa", required_argument, nullptr, kSamples},
    {"gpus", required_argument, nullptr, kGpus},
    {nullptr, 0, nullptr, 0},
};

// std::stoi semantics: optional whitespace/sign, longest decimal prefix, int range.
bool toInt(const char* text, int& out) {
    char* endp = nullptr;
    errno = 0;
    long v = std::strtol(text, &endp, 10);
    if (endp == text || errno == ERANGE || v < INT_MIN || v > INT_MAX) return false;
    out = (int)v;
    return true;
}
bool fail(const char* msg) {
    std::cerr << "Error: " << msg << std::endl;
    return false;
}
}  // namespace

bool Options::parseCommandLine(int argc, char* argv[]) {
    optind = 1;
    int opt;
    if (const char* g = std::getenv("AS2_GPUS")) {
        if (!toInt(g, gpus_) || gpus_ < 1 || gpus_ > 64) return fail("AS2_GPUS is invalid.");
    }
    while ((opt = getopt_long(argc, argv, "t:w:h:o:", kLongOptions, nullptr)) != -1) {
        switch (opt) {
            case 'o': outputFilename_ = optarg; break;
            case kIntersectionOnly: intersectionOnly_ = true; break;
            case kBruteForce: bruteForce_ = true; break;
            case 't':
                if (!toInt(optarg, renderThreadsCount_)) return fail("Thread count is invalid.");
                if (renderThreadsCount_ <= 0) return fail("Thread count must be positive.");
                break;
            case 'w':
            case 'h': {
                int& dest = (opt == 'w') ? renderWidth_ : renderHeight_;
                if (!toInt(optarg, dest)) return fail("Width and/or height is invalid.");
                if (dest <= 0) return fail("Width and/or height must be positive.");
                break;
            }
            case kBounceDepth:
                if (!toInt(optarg, bounceDepth_)) return fail("Bounce depth is invalid.");
                if (bounceDepth_ < 0) return fail("Bounce depth must be non-negative.");
                break;
            case kSamples:
                if (!toInt(optarg, samples_)) return fail("Sample count is invalid.");
                if (samples_ < 1 || samples_ > 16) return fail("Sample count must be between 1 and 16.");
                break;
            case kGpus:
                if (!toInt(optarg, gpus_)) return fail("GPU count is invalid.");
                if (gpus_ < 1 || gpus_ > 64) return fail("GPU count must be between 1 and 64.");
                break;
            case kHelp:
            case '?':
            default:
                printHelp(argv[0]);
                return false;
        }
    }
    for (; optind < argc; optind++) inputFilenames_.push_back(argv[optind]);
    if (inputFilenames_.empty()) return fail("At least one input file must be specified.");
    if (outputFilename_.empty()) return fail("An output file must be specified.");
    if (samples_ > 1 && intersectionOnly_) return fail("--aa cannot be combined with --intersection-only.");
    return true;
}

void Options::printHelp(const char* prog) {
    std::cerr << "Usage: " << prog << " [options] -o <output file> <input files>..." << std::endl;
}

}  // namespace as2
