// obj.cpp -- scene / material / camera text parsers with the grammar of the reference's Obj.hs.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

#include "squigly.hpp"

namespace squigly {

std::string readFile(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(path + ": openFile: does not exist (No such file or directory)");
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

namespace {

// A tiny recursive-descent reader over the same primitives Obj.hs builds from parsec.
class Reader {
  public:
    explicit Reader(const std::string &text) : s_(text) {}
    bool atEnd() const { return pos_ >= s_.size(); }
    char peek() const { return atEnd() ? '\0' : s_[pos_]; }
    void spaces() { while (!atEnd() && std::isspace((unsigned char)s_[pos_])) ++pos_; }          // parsec `spaces`
    bool string(const char *lit) {                                                               // `string`, no backtracking needed here
        size_t n = 0; while (lit[n]) ++n;
        if (s_.compare(pos_, n, lit) != 0) return false;
        pos_ += n; return true;
    }
    // word = many1 (noneOf " \t\n\r\f\v") <* spaces      (Obj.hs:131-132)
    bool word(std::string &out) {
        size_t b = pos_;
        while (!atEnd() && !std::isspace((unsigned char)s_[pos_])) ++pos_;
        out = s_.substr(b, pos_ - b); spaces();
        return !out.empty();
    }
    // fractional (Obj.hs:115-121): "-"? digit* "."? digit*  then `read`: correctly rounded decimal -> Float
    bool fractional(float &out) {
        size_t b = pos_, p = pos_;
        if (p < s_.size() && s_[p] == '-') ++p;
        while (p < s_.size() && std::isdigit((unsigned char)s_[p])) ++p;
        if (p < s_.size() && s_[p] == '.') ++p;
        while (p < s_.size() && std::isdigit((unsigned char)s_[p])) ++p;
        std::string tok = s_.substr(b, p - b);
        // Haskell `read :: String -> Float` accepts -?digits and -?digits.digits only: "1.", ".5", "-.5", "-" and "" are
        // "Prelude.read: no parse" (the reference crashes on them; here the parse fails)
        {
            size_t q = 0;
            if (q < tok.size() && tok[q] == '-') ++q;
            const size_t d1 = q;
            while (q < tok.size() && std::isdigit((unsigned char)tok[q])) ++q;
            if (q == d1) return false;
            if (q < tok.size()) {                 // tok[q] == '.'
                ++q;
                if (q >= tok.size()) return false;
            }
        }
        out = std::strtof(tok.c_str(), nullptr);
        pos_ = p;
        return true;
    }
    // vec3 (Obj.hs:166-171)
    bool vec3(V3 &v) {
        if (!fractional(v.x)) return false;
        spaces();
        if (!fractional(v.y)) return false;
        spaces();
        if (!fractional(v.z)) return false;
        spaces();
        return true;
    }
    bool natural(long &out) {                // many1 digit <* spaces
        if (!std::isdigit((unsigned char)peek())) return false;
        long v = 0;
        while (std::isdigit((unsigned char)peek())) { v = v * 10 + (peek() - '0'); ++pos_; }
        spaces(); out = v; return true;
    }
    void skip(size_t n = 1) { pos_ += n; }
    size_t pos() const { return pos_; }

  private:
    const std::string &s_;
    size_t pos_ = 0;
};

struct Face { long i1, i2, i3; };
struct Object { std::vector<V3> verts; std::string mtl; std::vector<Face> faces; };    // Obj.hs:90-94

[[noreturn]] void patternFail(const char *where) {
    // the reference crashes on `let Right ... = parse ...` (Obj.hs:51,53)
    throw std::runtime_error(std::string("Irrefutable pattern failed for pattern Right ") + where);
}

// loadObjFile = (,) <$> mtllib <*> many parseObj     (Obj.hs:96-107)
std::pair<std::string, std::vector<Object>> loadObjFile(const std::string &text) {
    Reader r(text);
    std::string lib;
    if (!r.string("mtllib")) patternFail("(mtllib', objs)");
    r.spaces();
    if (!r.word(lib)) patternFail("(mtllib', objs)");
    std::vector<Object> objs;
    while (r.peek() == 'o') {                         // objectName = char 'o' *> spaces *> many1 (alphaNum <|> oneOf "._") <* spaces
        r.skip(); r.spaces();
        bool any = false;
        while (std::isalnum((unsigned char)r.peek()) || r.peek() == '.' || r.peek() == '_') { r.skip(); any = true; }
        if (!any) patternFail("(mtllib', objs)");
        r.spaces();
        Object o;
        while (r.peek() == 'v') {                     // vertex = char 'v' *> spaces *> fmap swapYZ vec3  (Obj.hs:109-113)
            r.skip(); r.spaces();
            V3 p;
            if (!r.vec3(p)) patternFail("(mtllib', objs)");
            o.verts.push_back(V3{p.x, p.z, p.y});
        }
        if (!r.string("usemtl")) patternFail("(mtllib', objs)");      // materialName (Obj.hs:125-126)
        r.spaces();
        if (!r.word(o.mtl)) patternFail("(mtllib', objs)");
        // optional parseS (Obj.hs:134-135): try (string "s on") <|> string "s off".  parsec's `string` consumes what it
        // matched before it fails, so any other text starting with 's' is a parse error (the reference aborts), not a skip
        if (r.peek() == 's') {
            if (!(r.string("s on") || r.string("s off"))) patternFail("(mtllib', objs)");
            r.spaces();
        }
        while (r.peek() == 'f') {                     // face (Obj.hs:137-147)
            r.skip(); r.spaces();
            Face f;
            if (!r.natural(f.i1) || !r.natural(f.i2) || !r.natural(f.i3)) patternFail("(mtllib', objs)");
            o.faces.push_back(f);
        }
        objs.push_back(std::move(o));
    }
    // parsec's `parse` does not demand end of input and neither does the reference: whatever follows the last object is
    // ignored.  Say so, a truncated scene is otherwise rendered silently.
    if (!r.atEnd()) std::fprintf(stderr, "squigly: warning: .obj input ignored from offset %zu on (no further `o` object; Obj.hs:96-97 stops here too)\n", r.pos());
    return {lib, std::move(objs)};
}

// loadMtlFile = many loadMtl   (Obj.hs:146-161)
std::vector<std::pair<std::string, Material>> loadMtlFile(const std::string &text) {
    Reader r(text);
    std::vector<std::pair<std::string, Material>> mats;
    while (r.string("newmtl ")) {
        std::string name; Material m;
        if (!r.word(name)) patternFail("mats");
        r.spaces();
        if (!r.string("reflective ") || !r.fractional(m.reflective)) patternFail("mats");
        r.spaces();
        if (!r.vec3(m.surfColor)) patternFail("mats");
        r.spaces();
        if (!r.string("emissive ") || !r.fractional(m.emissive)) patternFail("mats");
        r.spaces();
        if (!r.vec3(m.emitColor)) patternFail("mats");
        r.spaces();
        mats.emplace_back(std::move(name), m);
    }
    return mats;
}

// c_ij = sum_k a_ik b_kj with `sum` = foldl (+) 0 (Data.Matrix multStd; package un-vendored, order unpinned)
void mul3(const float *a, const float *b, float *c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < 3; ++k) acc = acc + a[3 * i + k] * b[3 * k + j];
            c[3 * i + j] = acc;
        }
}

}  // namespace

// foldr1 (*) [Rz alp, Ry bet, Rx gam] = Rz * (Ry * Rx)     (Geometry.hs:90-102)
void rotMatrixRads(float alp, float bet, float gam, float out[9]) {
    const float rz[9] = {std::cos(alp), -std::sin(alp), 0, std::sin(alp), std::cos(alp), 0, 0, 0, 1};
    const float ry[9] = {std::cos(bet), 0, std::sin(bet), 0, 1, 0, -std::sin(bet), 0, std::cos(bet)};
    const float rx[9] = {1, 0, 0, 0, std::cos(gam), -std::sin(gam), 0, std::sin(gam), std::cos(gam)};
    float yx[9];
    mul3(ry, rx, yx);
    mul3(rz, yx, out);
}

ParsedScene trisFromObj(bool debug, const std::string &objContents, const std::string &dataDir) {
    auto parsed = loadObjFile(objContents);
    const std::string mtlFile = readFile(dataDir + parsed.first);          // Obj.hs:52
    auto mats = loadMtlFile(mtlFile);
    const auto &objs = parsed.second;
    // makeScene (Obj.hs:73-86): allVerts = concatMap verts objs ; one copy of the object per matching material
    std::vector<V3> allVerts;
    for (const auto &o : objs) allVerts.insert(allVerts.end(), o.verts.begin(), o.verts.end());
    ParsedScene out;
    for (const auto &m : mats) out.materials.push_back(m.second);
    for (const auto &o : objs)
        for (size_t mi = 0; mi < mats.size(); ++mi) {
            if (o.mtl != mats[mi].first) continue;
            for (const Face &f : o.faces) {
                auto at = [&](long i) -> const V3 & {                      // vs !! (a-1)
                    if (i < 1) throw std::runtime_error("Prelude.!!: negative index");
                    if ((size_t)i > allVerts.size()) throw std::runtime_error("Prelude.!!: index too large");
                    return allVerts[(size_t)i - 1];
                };
                out.triangles.push_back(Triangle{at(f.i1), at(f.i2), at(f.i3), (uint32_t)mi});
            }
        }
    if (debug) {                                                           // Obj.hs:55-57
        if (!objs.empty()) std::cout << "Object {verts = <" << objs[0].verts.size() << ">, mtl = \"" << objs[0].mtl
                                     << "\", faces = <" << objs[0].faces.size() << ">}\n";
        for (const auto &m : mats)
            std::cout << "(\"" << m.first << "\",Mat {reflective = " << m.second.reflective << ", emissive = " << m.second.emissive << "})\n";
    }
    return out;
}

// parseCamera = Camera <$> vec3 <*> (unpackRotMatrix <$> vec3)   (Obj.hs:67-70)
Camera loadCamera(const std::string &path) {
    const std::string text = readFile(path);
    Reader r(text);
    V3 pos, ang;
    if (!r.vec3(pos) || !r.vec3(ang)) throw std::runtime_error("Failed to parse /data/camera");   // Obj.hs:65
    Camera c;
    c.position = pos;
    rotMatrixRads(ang.x, ang.y, ang.z, c.rotation);
    return c;
}

}  // namespace squigly
