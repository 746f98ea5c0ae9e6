// parsers.h — `.rti` scene and `.obj` mesh readers: same grammar, warnings and error
// strings as the reference (src/parsers.h:5-49, src/parsers.cpp:5-374),
// re-implemented as a single-pass scanner over the file buffer (no per-line
// istringstream / per-token ostringstream), so million-triangle inputs parse at memory
// speed.
#pragma once
#include <string>
#include <vector>

#include "scene_model.h"

namespace as2 {

// One statement line cut into tokens with the reference's rules: whitespace separated,
// "..." quoted tokens, an unquoted token starting with '#' ends the line, and an empty
// token (e.g. "") ends the token list (src/parsers.cpp:24-76).
class LineLexer {
public:
    LineLexer(const char* begin, const char* end, int lineno) : p_(begin), end_(end), lineno_(lineno) {}
    // Returns false at end of line; otherwise sets [tb,te) to the token characters.
    bool next(const char*& tb, const char*& te);
    int lineno() const { return lineno_; }
private:
    const char* p_;
    const char* end_;
    int lineno_;
};

// std::stod semantics on a token: longest valid prefix, "invalid number" otherwise.
double parseNumber(const char* tb, const char* te, int lineno);

class RTIParser {
public:
    explicit RTIParser(Scene& scene) : scene_(scene) {}
    void parseFile(std::string filename);
private:
    void statement(const std::string& keyword, LineLexer& lex, const std::string& filename);
    Scene& scene_;
    Affine transform_ = Affine::identity();
    Material material_;
};

class OBJParser {
public:
    explicit OBJParser(Mesh& mesh) : mesh_(mesh) {}
    void parseFile(std::string filename);
private:
    Mesh& mesh_;
};

// Whole-file read; throws ParseException("file not found: ...") like the reference.
std::vector<char> slurpFile(const std::string& filename);

}  // namespace as2
