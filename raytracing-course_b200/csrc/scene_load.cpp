// scene_load.cpp -- hw5 scene text format (DIMENSIONS / SAMPLES / RAY_DEPTH / camera /
// NEW_PRIMITIVE blocks), host side.  Same observable behaviour as Scene::Load
// (src/sceneload.cpp:112-176) and LoadPrimitive (src/sceneload.cpp:35-110), written as a
// table-driven reader instead of two switch statements.
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

using PrimHandler = std::function<void(std::istream&, Primitive&)>;
using SceneHandler = std::function<void(std::istream&, HostScene&)>;

const std::unordered_map<std::string, PrimHandler>& prim_table() {
    static const std::unordered_map<std::string, PrimHandler> t = {
        {"ELLIPSOID", [](std::istream& in, Primitive& p) { vec3 r{0, 0, 0}; in >> r; p = shaped(PT_ELLIPSOID, r); }},
        {"PLANE", [](std::istream& in, Primitive& p) { vec3 n{0, 0, 0}; in >> n; p = shaped(PT_PLANE, n); }},
        {"BOX", [](std::istream& in, Primitive& p) { vec3 s{0, 0, 0}; in >> s; p = shaped(PT_BOX, s); }},
        {"TRIANGLE", [](std::istream& in, Primitive& p) {
             vec3 a{0, 0, 0}, b{0, 0, 0}, c{0, 0, 0};
             in >> a >> b >> c;
             p = shaped(PT_TRIANGLE, a, b, c);
         }},
        {"COLOR", [](std::istream& in, Primitive& p) { in >> p.col; }},
        {"POSITION", [](std::istream& in, Primitive& p) { in >> p.pos; }},
        {"ROTATION", [](std::istream& in, Primitive& p) { in >> p.rot; }},
        {"METALLIC", [](std::istream&, Primitive& p) { p.material = MAT_METALLIC; }},
        {"DIELECTRIC", [](std::istream&, Primitive& p) { p.material = MAT_DIELECTRIC; }},
        {"IOR", [](std::istream& in, Primitive& p) { in >> p.ior; }},
        {"EMISSION", [](std::istream& in, Primitive& p) { in >> p.emission; }},
    };
    return t;
}

const std::unordered_map<std::string, SceneHandler>& scene_table() {
    static const std::unordered_map<std::string, SceneHandler> t = {
        {"DIMENSIONS", [](std::istream& in, HostScene& s) { in >> s.cam.width >> s.cam.height; }},
        {"BG_COLOR", [](std::istream& in, HostScene& s) { in >> s.background; }},
        {"CAMERA_POSITION", [](std::istream& in, HostScene& s) { in >> s.cam.pos; }},
        {"CAMERA_RIGHT", [](std::istream& in, HostScene& s) { in >> s.cam.right; }},
        {"CAMERA_UP", [](std::istream& in, HostScene& s) { in >> s.cam.up; }},
        {"CAMERA_FORWARD", [](std::istream& in, HostScene& s) { in >> s.cam.forward; }},
        {"CAMERA_FOV_X", [](std::istream& in, HostScene& s) { in >> s.cam.fov_x; }},
        {"RAY_DEPTH", [](std::istream& in, HostScene& s) { in >> s.ray_depth; }},
        {"SAMPLES", [](std::istream& in, HostScene& s) { in >> s.samples; }},
    };
    return t;
}

// Reads one primitive block.  Returns the word that ended the block ("" for blank line / EOF).
std::string read_primitive(std::istream& in, Primitive& prim) {
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string word;
        ls >> word;
        if (word.empty()) return "";
        auto it = prim_table().find(word);
        if (it == prim_table().end()) return word;
        it->second(ls, prim);
    }
    return "";
}

}  // namespace

void HostScene::parse(const std::string& text) {
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string word;
        ls >> word;
        while (!word.empty()) {
            if (word == "NEW_PRIMITIVE") {
                Primitive prim;
                std::string leftover = read_primitive(in, prim);
                prim.orig = (int)prims.size();
                prims.push_back(prim);
                word = leftover;
                ls.clear();
                ls.str("");
                ls.setstate(std::ios::eofbit | std::ios::failbit);  // arguments of `leftover` are gone
                continue;
            }
            auto it = scene_table().find(word);
            if (it != scene_table().end()) it->second(ls, *this);
            else std::fprintf(stderr, "unexpected command(%s)\n", word.c_str());
            break;
        }
    }
}

}  // namespace rtc
