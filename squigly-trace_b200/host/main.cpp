// main.cpp -- the squigly-trace command line (app/Main.hs:13-75) driving the B200 backend.
// Flags follow the cmdargs record of the reference: -s/--samples, -d/--dimensions W,H, -p/--savepath,
// --objpath, -c/--camerapath, --debug, --debugpath, --cast ; plus --bounces, --gpus, --corrected, --seed.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>

#include "squigly.hpp"

using namespace squigly;

static void usage() {
    std::cout << "squigly-trace was made by Ruko (https://github.com/rukokarasu/)  [B200 backend]\n\n"
                 "squigly [OPTIONS]\n  A cute raytracer\n\nCommon flags:\n"
                 "  -s --samples=INT       How many samples per pixel to trace\n"
                 "  -d --dimensions=INT,INT Dimensions of the resulting image\n"
                 "  -p --savepath=FILE     Where to save the output\n"
                 "     --objpath=FILE      File to load .obj from\n"
                 "  -c --camerapath=FILE   File to load camera data from\n"
                 "     --debug             Run in debug mode\n"
                 "     --debugpath=FILE    File to write debug info to\n"
                 "     --cast              Raycast instead of raytracing (i.e. don't bounce rays)\n"
                 "     --bounces=INT       Intersections per path (reference: 3)\n"
                 "     --gpus=INT          Number of B200s to use\n"
                 "     --corrected         Width x height image with per-axis offsets (default: Lib.hs index convention)\n"
                 "     --seed=INT          RNG key\n"
                 "  -? --help              Display help message\n";
}

static std::string showTime() {      // formatTime defaultTimeLocale "%T%P UTC"
    std::time_t t = std::time(nullptr);
    char buf[64];
    std::strftime(buf, sizeof buf, "%T%P UTC", std::gmtime(&t));
    return buf;
}

int main(int argc, char **argv) {
    Settings st;
    try {
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i], val;
            auto eq = a.find('=');
            bool has_val = false;
            if (a.rfind("--", 0) == 0 && eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
            auto need = [&]() -> std::string {
                if (has_val) return val;
                if (i + 1 >= argc) throw std::runtime_error("Missing value for flag " + a);
                return argv[++i];
            };
            if (a == "-s" || a == "--samples") st.samples = std::stoi(need());
            else if (a == "-d" || a == "--dimensions") {
                const std::string d = need(); const auto c = d.find(',');
                if (c == std::string::npos) throw std::runtime_error("Could not parse dimensions, expected INT,INT");
                st.dimensions = {std::stoi(d.substr(0, c)), std::stoi(d.substr(c + 1))};
            } else if (a == "-p" || a == "--savepath") st.savePath = need();
            else if (a == "--objpath") st.objPath = need();
            else if (a == "-c" || a == "--camerapath") st.cameraPath = need();
            else if (a == "--debug") st.debug = true;
            else if (a == "--debugpath") st.debugPath = need();
            else if (a == "--cast") st.cast = true;
            else if (a == "--bounces") st.bounces = std::stoi(need());
            else if (a == "--gpus") st.gpus = std::stoi(need());
            else if (a == "--corrected") st.corrected = true;
            else if (a == "--seed") st.seed = std::stoull(need());
            else if (a == "-?" || a == "--help") { usage(); return 0; }
            else throw std::runtime_error("Unknown flag: " + a);
        }
        // main (Main.hs:35-47)
        const Camera cam = loadCamera(st.cameraPath);
        // loadBIH (Main.hs:63-75)
        ParsedScene tris = trisFromObj(st.debug, readFile(st.objPath));
        const BIH bih = makeBIH(std::move(tris));
        if (st.debug) {
            FILE *f = std::fopen(st.debugPath.c_str(), "w");
            if (!f) throw std::runtime_error(st.debugPath + ": openFile: does not exist (No such file or directory)");
            const std::string s = showBIH(bih); std::fwrite(s.data(), 1, s.size(), f); std::fclose(f);
            std::cout << "Wrote BIH to " << st.debugPath << "\n";
            std::cout << "BIH height is " << height(bih) << "\n";
            std::cout << "Length of longest leaf is " << longestLeaf(bih) << "\n";
            std::cout << "Number of leaves is " << numLeaves(bih) << "\n";
        }
        const SceneBIH scene = sceneFromBIH(bih, st.gpus);
        std::cout << "Rendering scene...\n";
        const auto t0 = std::chrono::steady_clock::now();
        std::cout << "Started at " << showTime() << "\n";
        const RenderReport rep = render(scene, cam, st);
        const auto t1 = std::chrono::steady_clock::now();
        std::cout << "Finished at " << showTime() << "\n";
        std::cout << "Took " << std::chrono::duration<double>(t1 - t0).count() << "s\n";
        if (st.debug)
            std::cout << "device " << rep.stats.device_ms << " ms, " << rep.stats.rays_traced << " rays, " << rep.stats.samples
                      << " samples, " << (rep.stats.rays_traced / (rep.stats.device_ms * 1e3)) << " Mrays/s\n";
    } catch (const std::exception &e) {
        std::cerr << "squigly-trace: " << e.what() << "\n";
        return 1;
    }
    return 0;
}
