// TEST INFRASTRUCTURE ONLY -- stand-in for cereal v1.2.2's <cereal/archives/portable_binary.hpp> (cereal is a network
// ExternalProject of the reference and is not vendored, vendor/cereal/CMakeLists.txt:10-11). It restates the byte layout
// of cereal's PortableBinaryOutputArchive / PortableBinaryInputArchive for exactly the types the reference serialises
// (main.cpp:147-166 -> lib/kdtree.h:228-230, :136-138, lib/triangle.h:89-92, src/serialize.h:10-24, lib/types.h:74-76):
//   * the archive's constructor writes ONE byte, 1 on a little-endian host (the flag the input archive uses to decide
//     about byte swapping);
//   * arithmetic values are written raw, little-endian, sizeof(T) bytes, no padding, no type or version tags;
//   * std::vector<T>: its size as a 64-bit unsigned integer (cereal's size_type), then the elements -- one contiguous
//     block for arithmetic T, element by element otherwise (same bytes either way);
//   * std::array<T, N>: the N elements, no size;
//   * a class: whatever its member serialize(Archive&) or the free serialize(Archive&, T&) found by ADL passes on,
//     in that order.
// WHICH members are written and in WHICH order is decided by the reference's own serialize() functions, compiled from
// /root/reference by oracle/build_ref.sh; only this primitive layer is a restatement ("parity unpinned" against a real
// cereal build, which cannot be made here).
#pragma once
#include <array>
#include <cstdint>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <type_traits>
#include <utility>
#include <vector>

namespace cereal {
namespace shim_detail {
template <class T, class A> auto has_member(int) -> decltype(std::declval<T&>().serialize(std::declval<A&>()), std::true_type());
template <class T, class A> std::false_type has_member(...);
} // namespace shim_detail

class PortableBinaryOutputArchive {
public:
    explicit PortableBinaryOutputArchive(std::ostream& s) : os_(s) {
        const std::uint8_t little = 1;
        os_.write(reinterpret_cast<const char*>(&little), 1);
    }
    template <class... T> void operator()(T&&... a) {
        int dummy[] = {0, (put(a), 0)...};
        (void)dummy;
    }

private:
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type put(const T& v) {
        os_.write(reinterpret_cast<const char*>(&v), sizeof(T));
    }
    template <class T, class Al> void put(std::vector<T, Al>& v) {
        put(static_cast<std::uint64_t>(v.size()));
        for (auto& e : v) put(e);
    }
    template <class T, std::size_t N> void put(std::array<T, N>& a) {
        for (auto& e : a) put(e);
    }
    template <class T>
    typename std::enable_if<!std::is_arithmetic<T>::value && decltype(shim_detail::has_member<T, PortableBinaryOutputArchive>(0))::value>::type
    put(T& t) {
        t.serialize(*this);
    }
    template <class T>
    typename std::enable_if<!std::is_arithmetic<T>::value && !decltype(shim_detail::has_member<T, PortableBinaryOutputArchive>(0))::value>::type
    put(T& t) {
        serialize(*this, t); // ADL: cereal::serialize (src/serialize.h) or ::serialize (lib/types.h:74-76)
    }
    std::ostream& os_;
};

class PortableBinaryInputArchive {
public:
    explicit PortableBinaryInputArchive(std::istream& s) : is_(s) {
        std::uint8_t little = 0;
        is_.read(reinterpret_cast<char*>(&little), 1);
        if (!is_ || little != 1) throw std::runtime_error("portable binary archive: not a little-endian stream");
    }
    template <class... T> void operator()(T&&... a) {
        int dummy[] = {0, (get(a), 0)...};
        (void)dummy;
    }

private:
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type get(T& v) {
        is_.read(reinterpret_cast<char*>(&v), sizeof(T));
        if (!is_) throw std::runtime_error("portable binary archive: truncated stream");
    }
    template <class T, class Al> void get(std::vector<T, Al>& v) {
        std::uint64_t n = 0;
        get(n);
        v.resize(static_cast<std::size_t>(n));
        for (auto& e : v) get(e);
    }
    template <class T, std::size_t N> void get(std::array<T, N>& a) {
        for (auto& e : a) get(e);
    }
    template <class T>
    typename std::enable_if<!std::is_arithmetic<T>::value && decltype(shim_detail::has_member<T, PortableBinaryInputArchive>(0))::value>::type
    get(T& t) {
        t.serialize(*this);
    }
    template <class T>
    typename std::enable_if<!std::is_arithmetic<T>::value && !decltype(shim_detail::has_member<T, PortableBinaryInputArchive>(0))::value>::type
    get(T& t) {
        serialize(*this, t);
    }
    std::istream& is_;
};
} // namespace cereal
