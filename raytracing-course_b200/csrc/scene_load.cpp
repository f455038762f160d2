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
#include <cstdio>
#include <functional>
#include <sstream>
#include <string>
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
    bool known(int dialect) const { return first <= dialect && dialect <= last; }
};

const std::unordered_map<std::string, Word<PrimFn>>& prim_table() {
    static const std::unordered_map<std::string, Word<PrimFn>> t = {
        {"ELLIPSOID", {[](std::istream& in, Primitive& p) { vec3 r{0, 0, 0}; in >> r; p = shaped(PT_ELLIPSOID, r); }, 1, 5}},
        {"PLANE", {[](std::istream& in, Primitive& p) { vec3 n{0, 0, 0}; in >> n; p = shaped(PT_PLANE, n); }, 1, 5}},
        {"BOX", {[](std::istream& in, Primitive& p) { vec3 s{0, 0, 0}; in >> s; p = shaped(PT_BOX, s); }, 1, 5}},
        {"TRIANGLE", {[](std::istream& in, Primitive& p) {
             vec3 a{0, 0, 0}, b{0, 0, 0}, c{0, 0, 0};
             in >> a >> b >> c;
             p = shaped(PT_TRIANGLE, a, b, c);
         }, 5, 5}},
        {"COLOR", {[](std::istream& in, Primitive& p) { in >> p.col; }, 1, 5}},
        {"POSITION", {[](std::istream& in, Primitive& p) { in >> p.pos; }, 1, 5}},
        {"ROTATION", {[](std::istream& in, Primitive& p) { in >> p.rot; }, 1, 5}},
        {"METALLIC", {[](std::istream&, Primitive& p) { p.material = MAT_METALLIC; }, 2, 5}},
        {"DIELECTRIC", {[](std::istream&, Primitive& p) { p.material = MAT_DIELECTRIC; }, 2, 5}},
        {"IOR", {[](std::istream& in, Primitive& p) { in >> p.ior; }, 2, 5}},
        {"EMISSION", {[](std::istream& in, Primitive& p) { in >> p.emission; }, 3, 5}},
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

// Reads one NEW_PRIMITIVE / NEW_LIGHT block.  Returns the word that ended the block ("" for blank line / EOF).
// `ls` is one line stream reused for every line of the file (constructing an istringstream per line was half of the
// time to read a 100k-triangle scene); clear() + str() give it exactly the state of a fresh one.
template <class Table, class Item>
std::string read_block(std::istream& in, std::istringstream& ls, const Table& table, int dialect, Item& item) {
    std::string line, word;
    while (std::getline(in, line)) {
        ls.clear();
        ls.str(line);
        word.clear();
        ls >> word;
        if (word.empty()) return "";
        auto it = table.find(word);
        if (it == table.end() || !it->second.known(dialect)) return word;
        it->second.fn(ls, item);
    }
    return "";
}

}  // namespace

void HostScene::parse(const std::string& text) {
    std::istringstream in(text);
    std::istringstream ls, block_ls;
    std::string line, word;
    while (std::getline(in, line)) {
        ls.clear();
        ls.str(line);
        word.clear();
        ls >> word;
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
