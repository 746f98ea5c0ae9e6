// main.cpp — the `as2` executable: same flow, messages and exit codes as the
// reference's main (src/main.cpp:40-85): options -> output-writability probe ->
// parse every positional .rti into ONE scene -> camera required -> render with a 4 Hz
// progress line -> PNG.  Only scene.renderScene() differs: it runs on the B200.
#include <signal.h>
#include <sys/time.h>

#include <cstdio>
#include <fstream>
#include <iostream>

#include "options.h"
#include "parsers.h"
#include "scene_model.h"
#include "writers.h"

using namespace as2;

static volatile sig_atomic_t g_tick = 0;

static void onAlarm(int) { g_tick = 1; }

static void showProgress(int complete, int total) {
    if (complete != total && !g_tick) return;
    std::printf("\rRendering scene (%d/%d) (%.1f%%) ...", complete, total, 100.0 * complete / total);
    std::fflush(stdout);
    if (complete == total) std::putchar('\n');
    g_tick = 0;
}

static void progressTimer(bool on) {
    if (on) signal(SIGALRM, onAlarm);
    struct itimerval itv = {};
    itv.it_value.tv_usec = itv.it_interval.tv_usec = on ? 250000 : 0;
    setitimer(ITIMER_REAL, &itv, nullptr);
    if (!on) signal(SIGALRM, SIG_DFL);
}

int main(int argc, char* argv[]) {
    if (!programOptions.parseCommandLine(argc, argv)) return 1;
    const std::string& out = programOptions.outputFilename_;
    {
        std::ofstream probe(out);
        if (!probe) {
            std::cerr << "Error: Output file is not writable." << std::endl;
            return 1;
        }
    }
    std::remove(out.c_str());

    Scene scene;
    for (const std::string& input : programOptions.inputFilenames_) {
        RTIParser parser(scene);
        try {
            parser.parseFile(input);
        } catch (const ParseException& e) {
            std::cerr << "Error: " << e.what() << std::endl;
            return 1;
        } catch (const MathException& e) {   // the reference terminates here; we report
            std::cerr << "Error: " << e.what() << std::endl;
            return 1;
        }
    }
    if (!scene.hasCamera()) {
        std::cerr << "Error: At least one camera must be specified." << std::endl;
        return 1;
    }

    progressTimer(true);
    Scene::RasterImage image(programOptions.renderHeight_, programOptions.renderWidth_);
    try {
        scene.renderScene(image, showProgress);
    } catch (const RenderException& e) {
        progressTimer(false);
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    progressTimer(false);

    try {
        PNGWriter(out).writeImage(image);
    } catch (const WriteException& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
