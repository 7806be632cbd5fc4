// main.cpp-style client of the REFERENCE's Encoder2 (encode/Encoder2.hpp:15-45) with the B200 engine registered in
// EncodingEngineCore2 (EngineCoreWithB200.cpp): grids and classifier wired like main.cpp:142-166, the reference's own
// job queue feeds the engines, the result is dumped in the golden format.
//   ref_core_main <luma.raw> <W> <H> <S> <T> <use_classifier> <nocpu>
#include "B200EncodingEngine.hpp"
#include "encode/Encoder2.hpp"

#include <cinttypes>
#include <cstdio>
#include <fstream>

int main(int argc, char** argv) {
    using namespace Frac2;
    if (argc < 8) return 2;
    const uint32_t W = std::atoi(argv[2]), H = std::atoi(argv[3]), S = std::atoi(argv[4]), T = std::atoi(argv[5]);
    std::vector<uint8_t> bytes((size_t)W * H);
    std::ifstream(argv[1], std::ios::binary).read(reinterpret_cast<char*>(bytes.data()), bytes.size());
    ImagePlane image(Size32u(W, H), W, std::move(bytes));
    Frac::encode_parameters_t params;
    params.sourceGridSize = S;
    params.targetGridSize = T;
    params.noclassifier = std::atoi(argv[6]) == 0;
    params.nocpu = std::atoi(argv[7]) != 0;
    std::unique_ptr<Classifier2> classifier = std::make_unique<BrightnessBlocksClassifier2>(image, image);
    if (params.noclassifier) classifier = std::make_unique<DummyClassifier>(image, image);
    auto classify = [&](const Point2du& origin, const Size32u& size) {
        UniformGridItem::ExtraData data;
        classifier->preclassify(origin, size, data);
        return data;
    };
    auto sourceGrid = createUniformGrid(image.size(), Size32u(S, S), Size32u(S / params.latticeSize, S / params.latticeSize), classify);
    auto targetGrid = createUniformGrid(image.size(), Size32u(T, T), Size32u(T, T), classify);
    try {
        Encoder2 encoder(image, params, sourceGrid, targetGrid, std::move(classifier), nullptr);
        const auto data = encoder.data();
        for (const auto& e : data.encoded) {
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
