// main.cpp-style client written ONLY against the reference's public encode/ API (same include paths,
// class names and call sequence as the reference's main.cpp:142-181), built against the drop-in headers
// in fractencode_b200/host.  Reads a raw u8 luma plane, encodes, decodes, prints the SURVEY-8c dump.
//   dropin_main <luma.raw> <W> <H> <S> <T> <noclassifier 0|1> <rms> <smax> <fma 0|1> [quadtree tmax tmin]
#include "encode/Encoder2.hpp"
#include "encode/Quantizer.hpp"
#include "image/Image2.hpp"
#include "image/partition2.hpp"

#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <fstream>
#include <map>

int main(int argc, char** argv) {
    using namespace Frac2;
    if (argc < 10) return 2;
    const uint32_t W = std::atoi(argv[2]), H = std::atoi(argv[3]);
    Frac::encode_parameters_t params;
    params.sourceGridSize = std::atoi(argv[4]);
    params.targetGridSize = std::atoi(argv[5]);
    params.noclassifier = std::atoi(argv[6]) != 0;
    params.rmsThreshold = std::atof(argv[7]);
    params.sMax = std::atof(argv[8]);
    params.fma = std::atoi(argv[9]) != 0;
    std::vector<uint8_t> bytes((size_t)W * H);
    std::ifstream(argv[1], std::ios::binary).read(reinterpret_cast<char*>(bytes.data()), bytes.size());
    try {
        ImagePlane image(Size32u(W, H), W, std::move(bytes));
        Frac::grid_encode_data_t data;
        if (argc >= 13 && std::string(argv[10]) == "quadtree") {
            QuadtreeEncoder2 enc(image, params, std::atoi(argv[11]), std::atoi(argv[12]));
            data = enc.data();
        } else {
            // ---- the reference's encode_image2 wiring (main.cpp:145-166) ----
            const Size32u gridSizeTarget(params.targetGridSize, params.targetGridSize);
            const Size32u gridSizeSource(params.sourceGridSize, params.sourceGridSize);
            const Size32u gridOffset = gridSizeSource / params.latticeSize;
            std::unique_ptr<Classifier2> classifier = std::make_unique<BrightnessBlocksClassifier2>(image, image);
            if (params.noclassifier) classifier = std::make_unique<DummyClassifier>(image, image);
            auto classifierCallback = [&](const Point2du& origin, const Size32u& size) {
                UniformGridItem::ExtraData d;
                classifier->preclassify(origin, size, d);
                return d;
            };
            auto sourceGrid = Frac2::createUniformGrid(image.size(), gridSizeSource, gridOffset, classifierCallback);
            auto targetGrid = Frac2::createUniformGrid(image.size(), gridSizeTarget, gridSizeTarget, classifierCallback);
            Encoder2 encoder(image, params, sourceGrid, targetGrid, std::move(classifier), nullptr);
            data = encoder.data();
        }
        // ---- decode (main.cpp:170-176) ----
        std::vector<uint8_t> imgData((size_t)W * H, 0);
        ImagePlane result({W, H}, W, std::move(imgData));
        Decoder2 decoder(result, -1, 0.00001, false, params.fma);
        const auto stats = decoder.decode(data);
        // ---- quantizer statistics (main.cpp:106-140) ----
        double max_c = -1, max_b = -1, min_c = 1.7976931348623157e308, min_b = 1.7976931348623157e308;
        for (const auto& d : data.encoded) {
            max_c = std::max(max_c, d.match.score.contrast); min_c = std::min(min_c, d.match.score.contrast);
            max_b = std::max(max_b, d.match.score.brightness); min_b = std::min(min_b, d.match.score.brightness);
        }
        std::map<Frac::Quantizerd::Int, int> cb, bb;
        if (max_c > min_c && max_b > min_b) {
            Frac::Quantizerd qb(min_b, max_b, 7), qc(min_c, max_c, 5);
            for (const auto& d : data.encoded) { ++cb[qc.quantized(d.match.score.contrast)]; ++bb[qb.quantized(d.match.score.brightness)]; }
        }
        std::sort(data.encoded.begin(), data.encoded.end(), [](const Frac::encode_item_t& a, const Frac::encode_item_t& b) {
            return std::make_tuple(a.y, a.x, a.w, a.h) < std::make_tuple(b.y, b.x, b.w, b.h);
        });
        for (const auto& e : data.encoded) {
            uint64_t d, s, o;
            std::memcpy(&d, &e.match.score.distance, 8); std::memcpy(&s, &e.match.score.contrast, 8); std::memcpy(&o, &e.match.score.brightness, 8);
            std::printf("%u %u %u %u | %u %u %u %u | t=%d d=%016" PRIx64 " s=%016" PRIx64 " o=%016" PRIx64 "\n", e.x, e.y, e.w, e.h, e.match.x, e.match.y,
                        e.match.sourceItemSize.x(), e.match.sourceItemSize.y(), (int)e.match.score.transform, d, s, o);
        }
        uint64_t fnv = 0xCBF29CE484222325ull;
        for (uint32_t y = 0; y < H; ++y)
            for (uint32_t x = 0; x < W; ++x) fnv = (fnv ^ result.value(x, y)) * 0x100000001B3ull;
        std::fprintf(stderr, "items=%zu decode_iterations=%d decode_rms=%.17g decode_fnv1a64=%016" PRIx64 " qbuckets=%zu,%zu\n", data.encoded.size(),
                     stats.iterations, stats.rms, fnv, cb.size(), bb.size());
    } catch (const std::exception& exc) {
        std::printf("EXCEPTION CAUGHT: %s\n", exc.what());
        return 1;
    }
    return 0;
}
