// frac_b200/core.hpp -- value types, image plane and grid partitions of the drop-in host API.
//
// Same names, members and meaning as the reference (sebsgit/fractencode) so that code written
// against its headers compiles unchanged:
//   Frac::Size<T>/Size32u          utils/size.hpp:8-70
//   Frac::Point2d<T>/Point2du      utils/point2d.hpp:8-53
//   Frac::TransformType            image/transform.h:16-25
//   Frac2::ImagePlane              image/Image2.hpp:79-152   (u8 pixels, explicit row stride)
//   Frac2::GridItemBase .. createUniformGrid   image/partition2.hpp:13-135
// Written from scratch for this project; storage is a std::vector instead of the reference's
// aligned malloc buffer (the pixels are copied to the GPU anyway).
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <iostream>
#include <type_traits>
#include <utility>
#include <vector>

#define FRAC_ASSERT(cond)                                                                          \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            std::cout << "Assert failed: " << __LINE__ << ' ' << __FILE__ << ' ' << #cond << '\n'; \
            std::exit(0); /* the reference exits with 0 on assertion failure (utils/Assert.hpp:4) */ \
        }                                                                                          \
    } while (0)

namespace Frac {

template <typename T, typename U> T convert(const U u) { return static_cast<T>(u); }

template <typename T> class Size {
public:
    constexpr Size() noexcept = default;
    constexpr Size(T x, T y) noexcept : _x(x), _y(y) {}
    T x() const noexcept { return _x; }
    T y() const noexcept { return _y; }
    void setX(const T& v) noexcept { _x = v; }
    void setY(const T& v) noexcept { _y = v; }
    bool operator==(const Size& o) const { return _x == o._x && _y == o._y; }
    bool operator!=(const Size& o) const { return !(*this == o); }
    Size operator/(const T& v) const { return Size(_x / v, _y / v); }
    Size operator*(const T& v) const { return Size(_x * v, _y * v); }
    bool isAligned(const T ax, const T ay) const { return _x % ax == 0 && _y % ay == 0; }
    Size align(const T ax, const T ay) const {
        const T rx = _x % ax, ry = _y % ay;
        return Size(_x + (rx ? ax - rx : 0), _y + (ry ? ay - ry : 0));
    }
    T area() const { return _x * _y; }
    friend std::ostream& operator<<(std::ostream& out, const Size& s) { return out << '{' << s.x() << 'x' << s.y() << '}'; }

private:
    T _x{}, _y{};
};
using Size32u = Size<uint32_t>;

template <typename T> class Point2d {
public:
    struct hash {
        constexpr auto operator()(const Point2d& p) const noexcept { return p._x ^ p._y; }
    };
    constexpr Point2d() noexcept = default;
    constexpr Point2d(const T x, const T y) noexcept : _x(x), _y(y) {}
    T x() const noexcept { return _x; }
    T y() const noexcept { return _y; }
    T& x() noexcept { return _x; }
    T& y() noexcept { return _y; }
    bool operator==(const Point2d& o) const noexcept { return o._x == _x && o._y == _y; }
    friend Point2d operator+(const Point2d& a, const Point2d& b) noexcept { return Point2d{a._x + b._x, a._y + b._y}; }
    friend std::ostream& operator<<(std::ostream& out, const Point2d& p) { return out << p.x() << ',' << p.y() << ' '; }

private:
    T _x{}, _y{};
};
using Point2du = Point2d<uint32_t>;

enum class TransformType { Id = 0, Rotate_90, Rotate_180, Rotate_270, Flip, Flip_Rotate_90, Flip_Rotate_180, Flip_Rotate_270 };

} // namespace Frac

namespace Frac2 {
using namespace Frac;

class ImagePlane {
public:
    ImagePlane(const ImagePlane&) = delete;
    ImagePlane& operator=(const ImagePlane&) = delete;
    ImagePlane() = default;
    ImagePlane(ImagePlane&&) = default;
    ImagePlane& operator=(ImagePlane&&) = default;
    ImagePlane(const Size32u& size, uint32_t stride) : _data((size_t)size.y() * stride), _stride(stride), _size(size) {}
    ImagePlane(const Size32u& size, uint32_t stride, std::initializer_list<uint8_t> init) : _data(init), _stride(stride), _size(size) {}
    ImagePlane(const Size32u& size, uint32_t stride, std::vector<uint8_t>&& init) : _data(std::move(init)), _stride(stride), _size(size) {}
    auto size() const noexcept { return _size; }
    auto width() const noexcept { return _size.x(); }
    auto height() const noexcept { return _size.y(); }
    auto stride() const noexcept { return _stride; }
    auto sizeInBytes() const noexcept { return _data.size(); }
    uint8_t* data() noexcept { return _data.data(); }
    const uint8_t* data() const noexcept { return _data.data(); }
    template <typename T = uint8_t> T value(int32_t x, int32_t y) const { return static_cast<T>(_data[(size_t)y * _stride + x]); }
    template <typename T, typename U> T value(const Point2d<U>& p) const { return value<T>(p.x(), p.y()); }
    void setValue(int32_t x, int32_t y, uint8_t v) { _data[(size_t)y * _stride + x] = v; }
    ImagePlane copy() const {
        ImagePlane r(_size, _stride);
        r._data = _data;
        return r;
    }

private:
    std::vector<uint8_t> _data;
    uint32_t _stride = 0;
    Size32u _size;
};

class GridItemBase {
public:
    Point2du origin;
    Size32u size;
    GridItemBase topLeft() const noexcept { return GridItemBase{origin, size / 2}; }
    GridItemBase topRight() const noexcept { return GridItemBase{origin + Point2du{size.x() / 2, 0}, size / 2}; }
    GridItemBase bottomLeft() const noexcept { return GridItemBase{origin + Point2du{0, size.y() / 2}, size / 2}; }
    GridItemBase bottomRight() const noexcept { return GridItemBase{origin + Point2du{size.x() / 2, size.y() / 2}, size / 2}; }
};
static_assert(sizeof(GridItemBase) == 4 * sizeof(uint32_t), "GridItemBase must stay 16 bytes (C-ABI fe_grid_item prefix)");

template <typename ExtraDataIn, bool isEmpty = std::is_empty_v<ExtraDataIn>> class GridItem;

template <typename ExtraDataIn> class GridItem<ExtraDataIn, false> : public GridItemBase {
public:
    using ExtraData = ExtraDataIn;
    ExtraData data;
    GridItem() noexcept : GridItemBase{Point2du{}, Size32u{}} {}
    GridItem(const Point2du& o, const Size32u& s) noexcept : GridItemBase{o, s} {}
    GridItem(const Point2du& o, const Size32u& s, ExtraData&& d) noexcept : GridItemBase{o, s}, data(std::move(d)) {}
};

template <typename ExtraDataIn> class GridItem<ExtraDataIn, true> : public GridItemBase {
public:
    using ExtraData = ExtraDataIn;
    GridItem() noexcept : GridItemBase{Point2du{}, Size32u{}} {}
    GridItem(const Point2du& o, const Size32u& s) noexcept : GridItemBase{o, s} {}
    GridItem(const Point2du& o, const Size32u& s, ExtraData&&) noexcept : GridItem(o, s) {}
};

template <typename Item> class GridPartition {
public:
    static GridPartition createEmpty(size_t n) {
        GridPartition r;
        r._items.resize(n);
        return r;
    }
    const auto& items() const noexcept { return _items; }
    void reserve(size_t n) { _items.reserve(n); }
    void add(const Point2du& o, const Size32u& s, typename Item::ExtraData&& d) { _items.push_back(Item{o, s, std::move(d)}); }
    void add(const Point2du& o, const Size32u& s) { _items.push_back(Item{o, s}); }

private:
    std::vector<Item> _items;
};

struct GridItemData {
    int32_t bb_classifierBin = -1;
};
using UniformGridItem = GridItem<GridItemData>;
using UniformGrid = GridPartition<UniformGridItem>;
static_assert(sizeof(UniformGridItem) == 20, "UniformGridItem must match fe_grid_item (20 bytes)");

// Row-major lattice, x fastest; item count per axis (extent - size) / step + 1.
template <typename Item = UniformGridItem>
GridPartition<Item> createUniformGrid(
    const Size32u& imageSize, const Size32u& itemSize, const Size32u& itemOffset,
    const std::function<typename Item::ExtraData(const Point2du&, const Size32u&)>& callback =
        [](const Point2du&, const Size32u&) -> typename Item::ExtraData { return typename Item::ExtraData{}; }) {
    FRAC_ASSERT(imageSize.isAligned(itemSize.x(), itemSize.y()) && "can't create grid partition on unaligned image!");
    FRAC_ASSERT(imageSize.isAligned(itemOffset.x(), itemOffset.y()) && "can't create grid partition with unaligned offset!");
    GridPartition<Item> result;
    for (uint32_t y = 0; y + itemSize.y() <= imageSize.y(); y += itemOffset.y())
        for (uint32_t x = 0; x + itemSize.x() <= imageSize.x(); x += itemOffset.x())
            result.add({x, y}, itemSize, callback({x, y}, itemSize));
    return result;
}

} // namespace Frac2
