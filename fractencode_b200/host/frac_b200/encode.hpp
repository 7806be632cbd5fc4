// frac_b200/encode.hpp -- the reference's encode/ API surface, backed by the B200 library.
//
// Same class names, constructor signatures, argument meaning and error behaviour as the
// reference (sebsgit/fractencode), so main.cpp-style code compiles unchanged and runs on the GPU:
//   Frac::transform_score_t / item_match_t / encode_item_t / grid_encode_data_t   encode/datatypes.h:8-26
//   Frac::encode_parameters_t                                                     encode/encode_parameters.h:5-14
//   Frac::Quantizer<T>                                                            encode/Quantizer.hpp:8-43
//   Frac::TransformMatcher                                                        encode/transformmatcher.h:18-150
//   Frac::copy                                                                    encode/DecodeUtils.hpp:9-25
//   Frac2::Classifier2 / DummyClassifier / BrightnessBlocksClassifier2            encode/Classifier2.hpp:10-66
//   Frac2::TransformEstimator2                                                    encode/TransformEstimator2.hpp:12-60
//   Frac2::ProgressReporter2 / StdoutReporter2 / AbstractEncodingEngine2 / EncodingEngineCore2
//                                                                                 encode/EncodingEngine2.hpp:13-180
//   Frac2::Encoder2 / Decoder2 / DummyReporter2                                   encode/Encoder2.hpp:9-105
// New (ours): Frac2::B200EncodingEngine2 (the engine the reference left as a TODO at
// encode/EncodingEngine2.cpp:21-26) and Frac2::QuadtreeEncoder2 (the reference parses --quadtree but
// has no quadtree).
//
// Everything numeric happens in libfractencode_b200.so through include/fractencode_b200.h; there
// is no CPU implementation behind these classes (errors surface as std::runtime_error, like the
// reference's engine-creation failures, encode/EncodingEngine2.cpp:27-29).
#pragma once

#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <memory>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>

#include "fractencode_b200.h"
#include "frac_b200/core.hpp"

namespace Frac {

struct transform_score_t {
    double distance = 100000.0;
    double contrast = 0.0;
    double brightness = 0.0;
    TransformType transform = TransformType::Id;
};
struct item_match_t {
    transform_score_t score;
    uint32_t x = 0;
    uint32_t y = 0;
    Size32u sourceItemSize;
};
struct encode_item_t {
    uint32_t x, y, w, h;
    item_match_t match;
};
struct grid_encode_data_t {
    std::vector<encode_item_t> encoded;
};
static_assert(sizeof(encode_item_t) == sizeof(fe_encode_item), "encode_item_t must match the C-ABI record (64 bytes)");

struct encode_parameters_t {
    int sourceGridSize = 16;
    int targetGridSize = 4;
    int latticeSize = 2;
    double rmsThreshold = 0.0;
    double sMax = -1.0;
    bool nogpu = false;
    bool nocpu = false;
    bool noclassifier = false;
    // B200 additions (defaults reproduce a reference built without FMA contraction)
    bool fma = false;      // brightness / decode evaluated with a fused multiply-add (reference -march=native build)
    int searchImpl = 0;    // fe_search_impl
};

template <typename T> class Quantizer {
    static_assert(std::is_arithmetic<T>::value, "cannot quantize non-arithmetic data");

public:
    using Int = uint64_t;
    explicit Quantizer(T minValue, T maxValue, int numberOfBits)
        : _min(minValue), _max(maxValue), _bits(numberOfBits), _step(std::abs(maxValue - minValue) / (1 << numberOfBits)),
          _maxQuantized((Int(1) << numberOfBits) - 1) {
        assert(maxValue > minValue);
        assert(numberOfBits > 1);
    }
    Int quantized(T value) const { return std::min(_maxQuantized, static_cast<Int>(std::floor((value - _min) / _step))); }
    T value(Int quant) const { return quant * _step + _min + _step / 2; }

private:
    const T _min, _max;
    const int _bits;
    const T _step;
    const Int _maxQuantized;
};
using Quantizerd = Quantizer<double>;

namespace b200 {

// One fe_ctx per host thread and device (the C ABI's threading rule), created on first use.
class Context {
public:
    explicit Context(int device = 0) {
        if (fe_create(&_ctx, device, nullptr) != FE_OK) throw std::runtime_error(std::string("fractencode_b200: ") + fe_last_error(nullptr));
    }
    ~Context() { fe_destroy(_ctx); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    fe_ctx* get() const noexcept { return _ctx; }
    void check(int rc) const {
        if (rc != FE_OK) throw std::runtime_error(std::string("fractencode_b200: ") + fe_last_error(_ctx));
    }
    static Context& threadLocal(int device = 0) {
        thread_local std::unique_ptr<Context> ctx[16];
        if (!ctx[device & 15]) ctx[device & 15] = std::make_unique<Context>(device);
        return *ctx[device & 15];
    }

private:
    fe_ctx* _ctx = nullptr;
};

inline fe_grid_item toAbi(const Frac2::UniformGridItem& it) {
    return fe_grid_item{it.origin.x(), it.origin.y(), it.size.x(), it.size.y(), it.data.bb_classifierBin};
}
inline std::vector<fe_grid_item> toAbi(const Frac2::UniformGrid& g) {
    static_assert(sizeof(Frac2::UniformGridItem) == sizeof(fe_grid_item), "layout");
    std::vector<fe_grid_item> v(g.items().size());
    if (!v.empty()) std::memcpy(v.data(), g.items().data(), v.size() * sizeof(fe_grid_item));
    return v;
}
inline encode_item_t fromAbi(const fe_encode_item& e) {
    encode_item_t r;
    r.x = e.x; r.y = e.y; r.w = e.w; r.h = e.h;
    r.match.score.distance = e.distance;
    r.match.score.contrast = e.contrast;
    r.match.score.brightness = e.brightness;
    r.match.score.transform = static_cast<TransformType>(e.transform);
    r.match.x = e.match_x; r.match.y = e.match_y;
    r.match.sourceItemSize = Size32u(e.src_w, e.src_h);
    return r;
}
inline void setImages(Context& c, const Frac2::ImagePlane& src, const Frac2::ImagePlane& tgt) {
    if (&src == &tgt || src.data() == tgt.data())
        c.check(fe_set_image(c.get(), src.data(), src.width(), src.height(), src.stride()));
    else
        c.check(fe_set_images(c.get(), src.data(), src.width(), src.height(), src.stride(), tgt.data(), tgt.width(), tgt.height(), tgt.stride()));
}

} // namespace b200

class TransformMatcher {
public:
    TransformMatcher(const double rmsThreshold, const double sMax) : _rmsThreshold(rmsThreshold), _sMax(sMax) {}
    double truncateSMax(const double s) const noexcept {
        if (_sMax > 0.0) return s > _sMax ? _sMax : (s < -_sMax ? -_sMax : s);
        return s;
    }
    bool checkDistance(const double d) const noexcept { return d <= _rmsThreshold; }
    double rmsThreshold() const noexcept { return _rmsThreshold; }
    double sMax() const noexcept { return _sMax; }
    // One domain against one range over the four rotations (reference: transformmatcher.h:38-46).
    transform_score_t match(const Frac2::ImagePlane& source, const Frac2::UniformGridItem& sourcePatch, const Frac2::ImagePlane& target,
                            const Frac2::UniformGridItem& targetPatch) const {
        auto& c = b200::Context::threadLocal();
        b200::setImages(c, source, target);
        fe_grid_item d = b200::toAbi(sourcePatch), r = b200::toAbi(targetPatch);
        d.bin = r.bin = -1;
        fe_params p{_rmsThreshold, _sMax, 0, 0, 0, 0};
        fe_encode_item out;
        c.check(fe_encode_level(c.get(), &d, 1, &r, 1, &p, &out));
        return b200::fromAbi(out).match.score;
    }

private:
    const double _rmsThreshold;
    const double _sMax;
};

// Frac::copy: target[item] = clamp(trunc(contrast * sample(source) + brightness)).
inline void copy(const Frac2::ImagePlane& source, Frac2::ImagePlane& target, const Frac2::GridItemBase& sourcePatch,
                 const Frac2::GridItemBase& targetPatch, const double contrast, const double brightness, TransformType transform) {
    auto& c = b200::Context::threadLocal();
    fe_encode_item e{};
    e.x = targetPatch.origin.x(); e.y = targetPatch.origin.y(); e.w = targetPatch.size.x(); e.h = targetPatch.size.y();
    e.contrast = contrast; e.brightness = brightness; e.transform = static_cast<int32_t>(transform);
    e.match_x = sourcePatch.origin.x(); e.match_y = sourcePatch.origin.y(); e.src_w = sourcePatch.size.x(); e.src_h = sourcePatch.size.y();
    FRAC_ASSERT(source.stride() == target.stride() && source.size() == target.size());
    c.check(fe_copy_items(c.get(), source.data(), target.data(), target.width(), target.height(), target.stride(), &e, 1, 0));
}

} // namespace Frac

namespace Frac2 {

class Classifier2 {
public:
    Classifier2(const ImagePlane& source, const ImagePlane& target) : _sourceImage(source), _targetImage(target) {}
    virtual ~Classifier2() = default;
    const ImagePlane& sourceImage() const noexcept { return _sourceImage; }
    const ImagePlane& targetImage() const noexcept { return _targetImage; }
    virtual bool compare(const UniformGridItem& source, const UniformGridItem& target) const = 0;
    virtual void preclassify(const Point2du&, const Size32u&, typename UniformGridItem::ExtraData&) const {}
    // B200 engine hook: does the search have to bucket blocks by brightness class?
    virtual bool usesBrightnessClasses() const noexcept { return false; }

protected:
    const ImagePlane& _sourceImage;
    const ImagePlane& _targetImage;
};

class DummyClassifier : public Classifier2 {
public:
    using Classifier2::Classifier2;
    bool compare(const UniformGridItem&, const UniformGridItem&) const override { return true; }
};

class BrightnessBlocksClassifier2 : public Classifier2 {
public:
    using Classifier2::Classifier2;
    // Class of one block (device computation of one item; batches go through fe_classify / the engine).
    static int getCategory(const ImagePlane& image, const UniformGridItem& item) {
        auto& c = Frac::b200::Context::threadLocal();
        c.check(fe_set_image(c.get(), image.data(), image.width(), image.height(), image.stride()));
        const fe_grid_item it = Frac::b200::toAbi(item);
        int32_t bin = -1;
        c.check(fe_classify(c.get(), 0, &it, 1, &bin));
        return bin;
    }
    bool compare(const UniformGridItem& a, const UniformGridItem& b) const override {
        int sa = a.data.bb_classifierBin, sb = b.data.bb_classifierBin;
        if (sa == -1) sa = getCategory(sourceImage(), a);
        if (sb == -1) sb = getCategory(targetImage(), b);
        return sa == sb;
    }
    // Leaves the bin at -1: the engine classifies every -1 block on the device in one batch, which is what
    // Classifier2::compare would compute lazily per pair (encode/Classifier2.cpp:70-81) -- same classes.
    void preclassify(const Point2du&, const Size32u&, typename UniformGridItem::ExtraData& data) const override { data.bb_classifierBin = -1; }
    bool usesBrightnessClasses() const noexcept override { return true; }
};

class TransformEstimator2 {
public:
    TransformEstimator2(const ImagePlane& sourceImage, const ImagePlane& targetImage, std::unique_ptr<Classifier2>&& classifier,
                        const std::shared_ptr<TransformMatcher>& matcher, const UniformGrid& sourceGrid)
        : _sourceImage(sourceImage), _targetImage(targetImage), _classifier(std::move(classifier)), _matcher(matcher), _sourceGrid(sourceGrid) {
        _rejectedMappings = 0;
    }
    // One range block against the whole source grid.  Batches should use estimateBatch (one device pass).
    item_match_t estimate(const UniformGridItem& targetItem) const {
        std::vector<UniformGridItem> one{targetItem};
        return estimateBatch(Frac::b200::Context::threadLocal(), one, Frac::encode_parameters_t{}).front().match;
    }
    std::vector<encode_item_t> estimateBatch(Frac::b200::Context& c, const std::vector<UniformGridItem>& targets,
                                             const Frac::encode_parameters_t& params) const {
        Frac::b200::setImages(c, _sourceImage, _targetImage);
        const std::vector<fe_grid_item> dom = Frac::b200::toAbi(_sourceGrid);
        std::vector<fe_grid_item> rng(targets.size());
        for (size_t i = 0; i < targets.size(); ++i) rng[i] = Frac::b200::toAbi(targets[i]);
        fe_params p{_matcher->rmsThreshold(), _matcher->sMax(), _classifier->usesBrightnessClasses() ? 1 : 0, params.fma ? 1 : 0,
                    params.searchImpl, 0};
        std::vector<fe_encode_item> out(targets.size());
        fe_stats before{}, after{};
        fe_get_stats(c.get(), &before);
        c.check(fe_encode_level(c.get(), dom.data(), dom.size(), rng.data(), rng.size(), &p, out.data()));
        fe_get_stats(c.get(), &after);
        std::vector<encode_item_t> res(out.size());
        for (size_t i = 0; i < out.size(); ++i) res[i] = Frac::b200::fromAbi(out[i]);
        // statistics only, like the reference's atomic counter: (domain, range) pairs the classifier excluded
        _rejectedMappings += (uint64_t)dom.size() * targets.size() - (after.matches - before.matches) / 4;
        return res;
    }
    uint64_t rejectedMappings() const noexcept { return _rejectedMappings; }
    const UniformGrid& sourceGrid() const noexcept { return _sourceGrid; }
    const TransformMatcher& matcher() const noexcept { return *_matcher; }

private:
    const ImagePlane& _sourceImage;
    const ImagePlane& _targetImage;
    std::unique_ptr<Classifier2> _classifier;
    std::shared_ptr<TransformMatcher> _matcher;
    const UniformGrid& _sourceGrid;
    mutable std::atomic<uint64_t> _rejectedMappings;
};

class ProgressReporter2 {
public:
    virtual ~ProgressReporter2() {}
    virtual void log(size_t done, size_t total) = 0;
};

class StdoutReporter2 : public ProgressReporter2 {
public:
    void log(size_t done, size_t total) override {
        const auto now = std::chrono::system_clock::now();
        if (std::chrono::duration<double>(now - _last).count() > 0.3) {
            _last = now;
            std::cout << '\r' << (100.0 * done) / total << std::flush;
        }
    }

private:
    std::chrono::system_clock::time_point _last;
};

class DummyReporter2 : public ProgressReporter2 {
public:
    void log(size_t, size_t) override {}
};

// The engine plug-in interface: init() once, encode(item) per job, finalize() once, then result().
class AbstractEncodingEngine2 {
public:
    AbstractEncodingEngine2(const encode_parameters_t& params, const ImagePlane& sourceImage, const UniformGrid& sourceGrid)
        : _parameters(params), _image(sourceImage), _source(sourceGrid) {}
    virtual ~AbstractEncodingEngine2() = default;
    virtual void encode(const UniformGridItem& targetItem) {
        _result.push_back(this->encode_impl(targetItem));
        ++_tasksDone;
    }
    void setName(const std::string& name) { _name = name; }
    const std::string& name() const noexcept { return _name; }
    std::vector<encode_item_t> result() const { return _result; }
    int tasksDone() const noexcept { return _tasksDone; }
    virtual void init() {}
    virtual void finalize() {}

protected:
    virtual encode_item_t encode_impl(const UniformGridItem& targetItem) const = 0;
    const encode_parameters_t _parameters;
    const ImagePlane& _image;
    const UniformGrid& _source;
    std::vector<encode_item_t> _result; // owned here so batching engines can fill it in finalize()
    std::string _name;
    int _tasksDone = 0;
};

// The GPU engine: encode() only queues the range block; finalize() searches the whole batch on the
// device (one fe_encode_level), so result() is complete when EncodingEngineCore2 reads it after join.
class B200EncodingEngine2 : public AbstractEncodingEngine2 {
public:
    B200EncodingEngine2(const encode_parameters_t& params, const ImagePlane& sourceImage, const UniformGrid& sourceGrid,
                        const TransformEstimator2& estimator, int device = 0)
        : AbstractEncodingEngine2(params, sourceImage, sourceGrid), _estimator(estimator), _device(device) {}
    void init() override { _ctx = std::make_unique<Frac::b200::Context>(_device); }
    void encode(const UniformGridItem& targetItem) override {
        _queue.push_back(targetItem);
        ++_tasksDone;
    }
    void finalize() override {
        if (_queue.empty()) return;
        const auto items = _estimator.estimateBatch(*_ctx, _queue, _parameters);
        _result.insert(_result.end(), items.begin(), items.end());
        _queue.clear();
    }

protected:
    encode_item_t encode_impl(const UniformGridItem& targetItem) const override {
        std::vector<UniformGridItem> one{targetItem};
        return _estimator.estimateBatch(*_ctx, one, _parameters).front();
    }

private:
    const TransformEstimator2& _estimator;
    const int _device;
    std::unique_ptr<Frac::b200::Context> _ctx;
    std::vector<UniformGridItem> _queue;
};

// Owner of the engines: one host thread per engine popping range blocks from a shared queue
// (reference: encode/EncodingEngine2.hpp:118-171, with the lost-wakeup wait replaced by join()).
class EncodingEngineCore2 {
public:
    EncodingEngineCore2(const encode_parameters_t& params, const ImagePlane& image, const UniformGrid& gridSource,
                        const TransformEstimator2& estimator, ProgressReporter2* reporter)
        : _estimator(estimator), _reporter(reporter) {
        FRAC_ASSERT(reporter);
        if (params.nogpu) throw std::runtime_error("fractencode_b200: nogpu requested but this build has no CPU engine");
        int devices = 1;
        if (const char* e = std::getenv("FRAC_B200_DEVICES")) devices = std::max(1, std::atoi(e));
        for (int d = 0; d < devices; ++d) {
            auto engine = std::make_unique<B200EncodingEngine2>(params, image, gridSource, _estimator, d);
            engine->setName("b200 " + std::to_string(d));
            _engines.push_back(std::move(engine));
        }
    }
    void encode(const UniformGrid& gridTarget) {
        const auto& jobs = gridTarget.items();
        std::mutex queueMutex;
        size_t next = 0;
        const size_t chunk = std::max<size_t>(1, (jobs.size() + _engines.size() - 1) / _engines.size());
        std::vector<std::thread> threads;
        std::vector<std::string> errors(_engines.size());
        for (size_t i = 0; i < _engines.size(); ++i) {
            threads.emplace_back([&, i]() {
                try {
                    _engines[i]->init();
                    for (;;) {
                        size_t b, e;
                        {
                            std::lock_guard<std::mutex> lock(queueMutex);
                            if (next >= jobs.size()) break;
                            b = next;
                            e = std::min(jobs.size(), b + chunk);
                            next = e;
                            _reporter->log(e, jobs.size());
                        }
                        for (size_t j = b; j < e; ++j) _engines[i]->encode(jobs[j]);
                    }
                    _engines[i]->finalize();
                } catch (const std::exception& exc) {
                    errors[i] = exc.what();
                }
            });
        }
        for (auto& t : threads) t.join();
        for (const auto& e : errors)
            if (!e.empty()) throw std::runtime_error(e);
        for (auto& engine : _engines) {
            const auto part = engine->result();
            _result.encoded.insert(_result.encoded.end(), part.begin(), part.end());
        }
    }
    const grid_encode_data_t result() const { return _result; }

private:
    std::vector<std::unique_ptr<AbstractEncodingEngine2>> _engines;
    const TransformEstimator2& _estimator;
    grid_encode_data_t _result;
    ProgressReporter2* _reporter; // not owned
};

class Encoder2 {
public:
    struct encode_stats_t {
        uint64_t rejectedMappings = 0;
        uint64_t totalMappings = 0;
        void print() {
            std::cout << "classifier rejected " << rejectedMappings << " out of " << totalMappings << " comparisons ("
                      << (100.0 * rejectedMappings) / totalMappings << ")%\n";
        }
    };
    Encoder2(const ImagePlane& image, const encode_parameters_t& p, const UniformGrid& sourcePartition, const UniformGrid& targetPartition,
             std::unique_ptr<Classifier2>&& classifier, std::unique_ptr<ProgressReporter2>&& reporter)
        : _encodeParameters(p), _reporter(reporter ? std::move(reporter) : std::make_unique<DummyReporter2>()) {
        _estimator.reset(new TransformEstimator2(image, image, std::move(classifier), std::make_shared<TransformMatcher>(p.rmsThreshold, p.sMax),
                                                 sourcePartition));
        _engine.reset(new EncodingEngineCore2(_encodeParameters, image, sourcePartition, *_estimator, _reporter.get()));
        _engine->encode(targetPartition);
        _stats.totalMappings = sourcePartition.items().size() * targetPartition.items().size();
    }
    grid_encode_data_t data() const { return _engine->result(); }

private:
    const encode_parameters_t _encodeParameters;
    mutable encode_stats_t _stats;
    std::unique_ptr<TransformEstimator2> _estimator;
    std::unique_ptr<EncodingEngineCore2> _engine;
    std::unique_ptr<ProgressReporter2> _reporter;
};

// Quadtree encoder (ours): level T block emitted when checkDistance(best) or T == tMin, else split into
// topLeft/topRight/bottomLeft/bottomRight; domains of level T = createUniformGrid(2T, step T).
class QuadtreeEncoder2 {
public:
    QuadtreeEncoder2(const ImagePlane& image, const encode_parameters_t& p, uint32_t tMax, uint32_t tMin, int device = 0) {
        Frac::b200::Context c(device);
        c.check(fe_set_image(c.get(), image.data(), image.width(), image.height(), image.stride()));
        fe_params fp{p.rmsThreshold, p.sMax, p.noclassifier ? 0 : 1, p.fma ? 1 : 0, p.searchImpl, 4};
        std::vector<fe_encode_item> out((size_t)(image.width() / tMin) * (image.height() / tMin));
        size_t n = 0;
        size_t counts[8] = {0};
        c.check(fe_encode_quadtree(c.get(), tMax, tMin, &fp, out.data(), out.size(), &n, counts));
        _data.encoded.resize(n);
        for (size_t i = 0; i < n; ++i) _data.encoded[i] = Frac::b200::fromAbi(out[i]);
        for (uint32_t T = tMax, l = 0; T >= tMin; T /= 2, ++l) _levelCounts.push_back(counts[l]);
    }
    grid_encode_data_t data() const { return _data; }
    const std::vector<size_t>& levelCounts() const noexcept { return _levelCounts; }

private:
    grid_encode_data_t _data;
    std::vector<size_t> _levelCounts;
};

class Decoder2 {
public:
    struct decode_stats_t {
        int iterations;
        double rms;
    };
    Decoder2(ImagePlane& target, const int nMaxIterations = -1, const double rmsEpsilon = 0.00001, bool saveDecodeSteps = false, bool fma = false)
        : _target(target), _iterations(nMaxIterations < 0 ? 300 : nMaxIterations), _rmsEpsilon(rmsEpsilon), _fma(fma) {
        (void)saveDecodeSteps; // decode_debugN.png dumps are file I/O, out of scope here
    }
    decode_stats_t decode(const grid_encode_data_t& data) {
        auto& c = Frac::b200::Context::threadLocal();
        static_assert(sizeof(encode_item_t) == sizeof(fe_encode_item), "layout");
        decode_stats_t st{0, 0.0};
        c.check(fe_decode(c.get(), reinterpret_cast<const fe_encode_item*>(data.encoded.data()), data.encoded.size(), _target.data(), _target.width(),
                          _target.height(), _target.stride(), _iterations, _rmsEpsilon, _fma ? 1 : 0, &st.iterations, &st.rms));
        return st;
    }

private:
    ImagePlane& _target;
    const int _iterations;
    const double _rmsEpsilon;
    const bool _fma;
};

} // namespace Frac2
