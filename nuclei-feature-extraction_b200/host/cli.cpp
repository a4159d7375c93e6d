// nfx-cli -- C++ host driver with the reference's argv contract (readme.md:10, src/args.rs:76-113):
//     nfx-cli [options] <input-geojson> <input-slide> <output-file> <feature-set>...
// main() mirrors src/main.rs:110-190: validate, load the geojson and the image, extract on the GPUs,
// write by extension. Errors print the reference's messages and exit 1 (src/args.rs:137-183).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <exception>

#include "nfx_host.hpp"

int main(int argc, char** argv) {
    using namespace nfxhost;
    try {
        const Args args = parse_args(argc, argv);
        if (args.verbose) std::fprintf(stderr, "Called Args : patch_size=%d batch_size=%d gpus=%zu sets=%zu\n", args.patch_size,
                                       args.batch_size, args.gpus.size(), args.feature_sets.size());
        const std::string ext = validate_paths(args);
        const bool timing = std::getenv("NFX_CLI_TIMING") != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            const auto t1 = std::chrono::steady_clock::now();
            if (timing) std::fprintf(stderr, "[nfx-cli] %-12s %.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
            t0 = t1;
        };
        feature_mask(args.feature_sets);                        // empty / duplicate sets fail before any work
        std::fprintf(stderr, "INFO Loading the geojson\n");      // main.rs:119
        const FeatureCollection geometry = load_geometry(args.geometry);
        lap("geojson");
        const Image image = load_input_image(args.slide);
        lap("image");
        std::fprintf(stderr, "INFO Extracting features\n");      // main.rs:143
        if (ext == "csv" && !args.via_trait && !args.host_csv) {
            extract_to_csv(geometry, image, args, args.output);   // rows formatted on each GPU, no DataFrame in between
            lap("extract+csv");
            return 0;
        }
        const DataFrame df = args.via_trait ? extract_via_trait(geometry, image, args) : extract(geometry, image, args);
        write_output(args.output, ext, df, args.host_csv || args.gpus.empty() ? -1 : args.gpus[0]);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ERROR %s\n", e.what());
        return 1;
    }
}
