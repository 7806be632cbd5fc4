// The registration a reference maintainer adds: EncodingEngineCore2's constructor (encode/EncodingEngine2.cpp:8-30) with the
// GPU-engine slot of lines 21-26 filled by the B200 engine.  This translation unit REPLACES encode/EncodingEngine2.cpp in the
// link; everything else -- EncodingEngineCore2::encode with its job queue and one thread per engine, Encoder2,
// TransformEstimator2, the CPU engines -- is the reference's own unmodified code from its headers.
#include "B200EncodingEngine.hpp"
#include "utils/Assert.hpp"

namespace Frac2 {

EncodingEngineCore2::EncodingEngineCore2(const encode_parameters_t& params, const ImagePlane& image, const UniformGrid& gridSource,
                                         const TransformEstimator2& estimator, ProgressReporter2* reporter)
    : _estimator(estimator), _reporter(reporter) {
    FRAC_ASSERT(reporter);
    auto add = [this](std::unique_ptr<AbstractEncodingEngine2> engine, const std::string& name) {
        engine->setName(name);
        _engines.push_back(std::move(engine));
    };
    // CPU engines exactly as upstream: one per hardware thread unless --nocpu
    for (unsigned i = 0; !params.nocpu && i < std::thread::hardware_concurrency(); ++i)
        add(std::make_unique<CpuEncodingEngine2>(params, image, gridSource, _estimator), "cpu " + std::to_string(i));
    // the slot upstream reserved for "CUDA, OpenCL engines"
    if (!params.nogpu) {
        try {
            add(std::make_unique<B200EncodingEngine>(params, image, gridSource), "B200");
        } catch (const std::exception& exc) {
            std::cout << "failed to create engine: " << exc.what();
        }
    }
}

} // namespace Frac2
