// Drives the reference's OWN types (ImagePlane, createUniformGrid, UniformGridItem, encode_item_t) through the
// B200EncodingEngine stub of INTEGRATION.md: init() -> encode(item)... -> finalize() -> result().
//   ref_engine_main <luma.raw> <W> <H> <S> <T> <use_classifier>
#include "B200EncodingEngine.hpp"

#include <cinttypes>
#include <cstdio>
#include <fstream>

int main(int argc, char** argv) {
    using namespace Frac2;
    if (argc < 7) return 2;
    const uint32_t W = std::atoi(argv[2]), H = std::atoi(argv[3]), S = std::atoi(argv[4]), T = std::atoi(argv[5]);
    std::vector<uint8_t> bytes((size_t)W * H);
    std::ifstream(argv[1], std::ios::binary).read(reinterpret_cast<char*>(bytes.data()), bytes.size());
    ImagePlane image(Size32u(W, H), W, std::move(bytes));
    Frac::encode_parameters_t params;
    params.sourceGridSize = S;
    params.targetGridSize = T;
    params.noclassifier = std::atoi(argv[6]) == 0;
    auto sourceGrid = createUniformGrid(image.size(), Size32u(S, S), Size32u(S / 2, S / 2));
    auto targetGrid = createUniformGrid(image.size(), Size32u(T, T), Size32u(T, T));
    try {
        B200EncodingEngine engine(params, image, sourceGrid);
        engine.setName("B200");
        engine.init();
        for (const auto& item : targetGrid.items()) engine.encode(item);
        engine.finalize();
        for (const auto& e : engine.result()) {
            uint64_t d, s, o;
            std::memcpy(&d, &e.match.score.distance, 8); std::memcpy(&s, &e.match.score.contrast, 8); std::memcpy(&o, &e.match.score.brightness, 8);
            std::printf("%u %u %u %u | %u %u %u %u | t=%d d=%016" PRIx64 " s=%016" PRIx64 " o=%016" PRIx64 "\n", e.x, e.y, e.w, e.h, e.match.x, e.match.y,
                        e.match.sourceItemSize.x(), e.match.sourceItemSize.y(), (int)e.match.score.transform, d, s, o);
        }
    } catch (const std::exception& exc) {
        std::printf("EXCEPTION CAUGHT: %s\n", exc.what());
        return 1;
    }
    return 0;
}
