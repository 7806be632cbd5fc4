// The reference-side binding of INTEGRATION.md section 2, kept compilable: this header includes the REFERENCE's own
// encode/EncodingEngine2.hpp (found through -I <reference checkout>) and our C ABI, nothing else.
#pragma once
#include "encode/EncodingEngine2.hpp"
#include "fractencode_b200.h"

#include <stdexcept>

#ifndef FRAC_FMA_BUILD
#if defined(__FMA__)
#define FRAC_FMA_BUILD 1
#else
#define FRAC_FMA_BUILD 0
#endif
#endif

namespace Frac2 {
class B200EncodingEngine : public AbstractEncodingEngine2 {
public:
    // same constructor shape as the OpenCLEncodingEngine the reference left commented out (encode/EncodingEngine2.cpp:23);
    // the classifier choice travels in encode_parameters_t::noclassifier like it does for main.cpp:152-154
    B200EncodingEngine(const encode_parameters_t& p, const ImagePlane& image, const UniformGrid& source, int device = 0)
        : AbstractEncodingEngine2(p, image, source), _useClassifier(!p.noclassifier), _device(device) {}
    ~B200EncodingEngine() override { fe_destroy(_ctx); }

    void init() override {
        if (fe_create(&_ctx, _device, nullptr) != FE_OK) throw std::runtime_error(fe_last_error(nullptr));
        check(fe_set_image(_ctx, _image.data(), _image.width(), _image.height(), _image.stride()));
    }
    void encode(const UniformGridItem& item) override { _queue.push_back(item); }
    void finalize() override {
        static_assert(sizeof(UniformGridItem) == sizeof(fe_grid_item), "20-byte layout");
        static_assert(sizeof(encode_item_t) == sizeof(fe_encode_item), "64-byte layout");
        fe_params fp{_parameters.rmsThreshold, _parameters.sMax, _useClassifier ? 1 : 0, FRAC_FMA_BUILD, FE_SEARCH_AUTO, 4};
        _batch.resize(_queue.size());
        check(fe_encode_level(_ctx, reinterpret_cast<const fe_grid_item*>(_source.items().data()), _source.items().size(),
                              reinterpret_cast<const fe_grid_item*>(_queue.data()), _queue.size(), &fp,
                              reinterpret_cast<fe_encode_item*>(_batch.data())));
        for (_next = 0; _next < _queue.size(); ++_next) AbstractEncodingEngine2::encode(_queue[_next]);
    }

protected:
    encode_item_t encode_impl(const UniformGridItem&) const override { return _batch[_next]; }

private:
    void check(int rc) const {
        if (rc != FE_OK) throw std::runtime_error(fe_last_error(_ctx));
    }
    const bool _useClassifier;
    const int _device;
    fe_ctx* _ctx = nullptr;
    std::vector<UniformGridItem> _queue;
    std::vector<encode_item_t> _batch;
    size_t _next = 0;
};
} // namespace Frac2
