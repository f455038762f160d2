// scene_load.cpp -- hw5 scene text format (DIMENSIONS / SAMPLES / RAY_DEPTH / camera /
// NEW_PRIMITIVE blocks), host side.  Same observable behaviour as Scene::Load
// (src/sceneload.cpp:112-176) and LoadPrimitive (src/sceneload.cpp:35-110), written as a
// table-driven reader instead of two switch statements.
//
// The earlier homework dialects (hw1..hw4 src/scene.cpp get_command / LoadPrimitive / LoadLight /
// Scene::Load) are the same reader with a smaller or different vocabulary: every table entry names
// the dialects [first, last] that know the word; hw2 adds AMBIENT_LIGHT and NEW_LIGHT blocks.
//
// Behaviours kept on purpose (scenes in the wild rely on them):
//  * the first word of a line is the command; arguments are read with operator>> semantics;
//  * a primitive block ends at the first empty line, at EOF, or at the first word that is not a
//    primitive attribute; that word is then handled as a scene-level command, but its
//    arguments are NOT available any more (the reference re-dispatches on the exhausted stream
//    of the NEW_PRIMITIVE line) -- in practice only a following NEW_PRIMITIVE matters;
//  * a shape line (PLANE/BOX/ELLIPSOID/TRIANGLE) resets every attribute given before it;
//  * unknown scene-level words are reported on stderr and skipped.
#include <charconv>
#include <cstdio>
#include <functional>
#include <sstream>
#include <string>
#include <string_view>
#include <unordered_map>

#include "scene_host.h"

namespace rtc {
namespace {

std::istream& operator>>(std::istream& in, vec3& v) { return in >> v.x >> v.y >> v.z; }
std::istream& operator>>(std::istream& in, quat& q) { return in >> q.x >> q.y >> q.z >> q.w; }

Primitive shaped(int type, vec3 a, vec3 b = {0, 0, 0}, vec3 c = {0, 0, 0}) {
    Primitive p;
    p.type = type;
    p.d0 = a;
    p.d1 = b;
    p.d2 = c;
    return p;
}

using PrimFn = std::function<void(std::istream&, Primitive&)>;
using SceneFn = std::function<void(std::istream&, HostScene&)>;
using LightFn = std::function<void(std::istream&, PointLight&)>;
template <class F>
struct Word {
    F fn;
    int first, last;  // dialects that know this word
    // Fast form of a primitive attribute (nearly every line of a 100k-triangle scene is one): the word takes `nfloat`
    // floats; when the line holds that many PLAIN decimal numbers (plain_float below) they are converted with
    // std::from_chars and handed to fn through a stream-free twin.  Anything else -- a missing, malformed or
    // out-of-range argument -- goes through operator>> as before, so that the partial-failure behaviour of the
    // reference's reader (sceneload.cpp uses operator>> throughout) is kept by construction.
    int nfloat = -1;
    void (*fast)(const float*, Primitive&) = nullptr;
    bool known(int dialect) const { return first <= dialect && dialect <= last; }
};

// [+-]digits[.digits][(e|E)[+-]digits] with at least one digit in the mantissa: the subset of what operator>>(float&)
// accepts (libstdc++ num_get -> strtof) on which std::from_chars gives the same, correctly rounded, value
bool plain_float(const char* b, const char* e, float& out) {
    const char* p = b;
    if (p < e && (*p == '+' || *p == '-')) ++p;
    int nd = 0;
    while (p < e && *p >= '0' && *p <= '9') { ++p; ++nd; }
    if (p < e && *p == '.') { ++p; while (p < e && *p >= '0' && *p <= '9') { ++p; ++nd; } }
    if (nd == 0) return false;
    if (p < e && (*p == 'e' || *p == 'E')) {
        ++p;
        if (p < e && (*p == '+' || *p == '-')) ++p;
        int ne = 0;
        while (p < e && *p >= '0' && *p <= '9') { ++p; ++ne; }
        if (ne == 0) return false;
    }
    if (p != e) return false;
    if (*b == '+') ++b;   // from_chars takes no leading plus
    const std::from_chars_result r = std::from_chars(b, e, out);
    return r.ec == std::errc() && r.ptr == e;   // over- / underflow: let operator>> decide
}
bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }   // the classic locale's
// the first `n` whitespace-separated tokens after `p` as plain floats
bool plain_floats(const char* p, const char* end, int n, float* out) {
    for (int i = 0; i < n; ++i) {
        while (p < end && is_space(*p)) ++p;
        const char* b = p;
        while (p < end && !is_space(*p)) ++p;
        if (b == p || !plain_float(b, p, out[i])) return false;
    }
    return true;
}

const std::unordered_map<std::string, Word<PrimFn>>& prim_table() {
    static const std::unordered_map<std::string, Word<PrimFn>> t = {
        {"ELLIPSOID", {[](std::istream& in, Primitive& p) { vec3 r{0, 0, 0}; in >> r; p = shaped(PT_ELLIPSOID, r); }, 1, 5,
                       3, [](const float* f, Primitive& p) { p = shaped(PT_ELLIPSOID, vec3{f[0], f[1], f[2]}); }}},
        {"PLANE", {[](std::istream& in, Primitive& p) { vec3 n{0, 0, 0}; in >> n; p = shaped(PT_PLANE, n); }, 1, 5,
                   3, [](const float* f, Primitive& p) { p = shaped(PT_PLANE, vec3{f[0], f[1], f[2]}); }}},
        {"BOX", {[](std::istream& in, Primitive& p) { vec3 s{0, 0, 0}; in >> s; p = shaped(PT_BOX, s); }, 1, 5,
                 3, [](const float* f, Primitive& p) { p = shaped(PT_BOX, vec3{f[0], f[1], f[2]}); }}},
        {"TRIANGLE", {[](std::istream& in, Primitive& p) {
             vec3 a{0, 0, 0}, b{0, 0, 0}, c{0, 0, 0};
             in >> a >> b >> c;
             p = shaped(PT_TRIANGLE, a, b, c);
         }, 5, 5,
         9, [](const float* f, Primitive& p) { p = shaped(PT_TRIANGLE, vec3{f[0], f[1], f[2]}, vec3{f[3], f[4], f[5]}, vec3{f[6], f[7], f[8]}); }}},
        {"COLOR", {[](std::istream& in, Primitive& p) { in >> p.col; }, 1, 5, 3, [](const float* f, Primitive& p) { p.col = vec3{f[0], f[1], f[2]}; }}},
        {"POSITION", {[](std::istream& in, Primitive& p) { in >> p.pos; }, 1, 5, 3, [](const float* f, Primitive& p) { p.pos = vec3{f[0], f[1], f[2]}; }}},
        {"ROTATION", {[](std::istream& in, Primitive& p) { in >> p.rot; }, 1, 5,
                      4, [](const float* f, Primitive& p) { p.rot.x = f[0]; p.rot.y = f[1]; p.rot.z = f[2]; p.rot.w = f[3]; }}},
        {"METALLIC", {[](std::istream&, Primitive& p) { p.material = MAT_METALLIC; }, 2, 5}},
        {"DIELECTRIC", {[](std::istream&, Primitive& p) { p.material = MAT_DIELECTRIC; }, 2, 5}},
        {"IOR", {[](std::istream& in, Primitive& p) { in >> p.ior; }, 2, 5, 1, [](const float* f, Primitive& p) { p.ior = f[0]; }}},
        {"EMISSION", {[](std::istream& in, Primitive& p) { in >> p.emission; }, 3, 5, 3, [](const float* f, Primitive& p) { p.emission = vec3{f[0], f[1], f[2]}; }}},
    };
    return t;
}

const std::unordered_map<std::string, Word<SceneFn>>& scene_table() {
    static const std::unordered_map<std::string, Word<SceneFn>> t = {
        {"DIMENSIONS", {[](std::istream& in, HostScene& s) { in >> s.cam.width >> s.cam.height; }, 1, 5}},
        {"BG_COLOR", {[](std::istream& in, HostScene& s) { in >> s.background; }, 1, 5}},
        {"CAMERA_POSITION", {[](std::istream& in, HostScene& s) { in >> s.cam.pos; }, 1, 5}},
        {"CAMERA_RIGHT", {[](std::istream& in, HostScene& s) { in >> s.cam.right; }, 1, 5}},
        {"CAMERA_UP", {[](std::istream& in, HostScene& s) { in >> s.cam.up; }, 1, 5}},
        {"CAMERA_FORWARD", {[](std::istream& in, HostScene& s) { in >> s.cam.forward; }, 1, 5}},
        {"CAMERA_FOV_X", {[](std::istream& in, HostScene& s) { in >> s.cam.fov_x; }, 1, 5}},
        {"RAY_DEPTH", {[](std::istream& in, HostScene& s) { in >> s.ray_depth; }, 2, 5}},
        {"SAMPLES", {[](std::istream& in, HostScene& s) { in >> s.samples; }, 3, 5}},
        {"AMBIENT_LIGHT", {[](std::istream& in, HostScene& s) { in >> s.ambient; }, 2, 2}},
    };
    return t;
}

// hw2 LoadLight (hw2/src/scene.cpp:120-168)
const std::unordered_map<std::string, Word<LightFn>>& light_table() {
    static const std::unordered_map<std::string, Word<LightFn>> t = {
        {"LIGHT_INTENSITY", {[](std::istream& in, PointLight& l) { in >> l.intensity; }, 2, 2}},
        {"LIGHT_POSITION", {[](std::istream& in, PointLight& l) { in >> l.pos; }, 2, 2}},
        {"LIGHT_DIRECTION", {[](std::istream& in, PointLight& l) { in >> l.dir; l.directed = 1; }, 2, 2}},
        {"LIGHT_ATTENUATION", {[](std::istream& in, PointLight& l) { in >> l.att; }, 2, 2}},
    };
    return t;
}

// std::getline over the scene text without copying it: lines end at '\n', the last one may lack it
struct Lines {
    const char* p;
    const char* end;
    bool next(std::string_view& line) {
        if (p >= end) return false;
        const char* e = p;
        while (e < end && *e != '\n') ++e;
        line = std::string_view(p, (size_t)(e - p));
        p = e < end ? e + 1 : e;
        return true;
    }
};
// the first word of a line, as `ls >> word` reads it
std::string_view first_word(std::string_view line, const char*& rest) {
    const char* p = line.data();
    const char* end = p + line.size();
    while (p < end && is_space(*p)) ++p;
    const char* b = p;
    while (p < end && !is_space(*p)) ++p;
    rest = p;
    return std::string_view(b, (size_t)(p - b));
}
template <class Item>
bool try_fast(const Word<std::function<void(std::istream&, Item&)>>&, const char*, const char*, Item&) { return false; }
bool try_fast(const Word<PrimFn>& w, const char* rest, const char* end, Primitive& item) {
    if (!w.fast) return false;
    float f[9];
    if (!plain_floats(rest, end, w.nfloat, f)) return false;
    w.fast(f, item);
    return true;
}

// Reads one NEW_PRIMITIVE / NEW_LIGHT block.  Returns the word that ended the block ("" for blank line / EOF).
// `ls` is one line stream reused for every line that needs operator>> (constructing an istringstream per line was half
// of the time to read a 100k-triangle scene; a line of plain numbers needs none: try_fast); clear() + str() give it
// exactly the state of a fresh one.
template <class Table, class Item>
std::string read_block(Lines& in, std::istringstream& ls, const Table& table, int dialect, Item& item) {
    std::string_view line;
    std::string key;
    while (in.next(line)) {
        const char* rest;
        const std::string_view word = first_word(line, rest);
        if (word.empty()) return "";
        key.assign(word.data(), word.size());
        auto it = table.find(key);
        if (it == table.end() || !it->second.known(dialect)) return key;
        if (try_fast(it->second, rest, line.data() + line.size(), item)) continue;
        ls.clear();
        ls.str(std::string(rest, (size_t)(line.data() + line.size() - rest)));   // the stream right after `ls >> word`
        it->second.fn(ls, item);
    }
    return "";
}

}  // namespace

void HostScene::parse(const std::string& text) {
    Lines in{text.data(), text.data() + text.size()};
    std::istringstream ls, block_ls;
    std::string_view line;
    std::string word;
    while (in.next(line)) {
        const char* rest;
        const std::string_view w = first_word(line, rest);
        word.assign(w.data(), w.size());
        ls.clear();
        ls.str(std::string(rest, (size_t)(line.data() + line.size() - rest)));
        while (!word.empty()) {
            const bool new_prim = word == "NEW_PRIMITIVE", new_light = word == "NEW_LIGHT" && dialect == DIALECT_HW2;
            if (new_prim || new_light) {
                std::string leftover;
                if (new_prim) {
                    Primitive prim;
                    leftover = read_block(in, block_ls, prim_table(), dialect, prim);
                    prim.orig = (int)prims.size();
                    prims.push_back(prim);
                } else {
                    PointLight light;
                    leftover = read_block(in, block_ls, light_table(), dialect, light);
                    point_lights.push_back(light);
                }
                word = leftover;
                ls.clear();
                ls.str("");
                ls.setstate(std::ios::eofbit | std::ios::failbit);  // arguments of `leftover` are gone
                continue;
            }
            auto it = scene_table().find(word);
            if (it != scene_table().end() && it->second.known(dialect)) it->second.fn(ls, *this);
            else std::fprintf(stderr, "unexpected command(%s)\n", word.c_str());
            break;
        }
    }
}

}  // namespace rtc
