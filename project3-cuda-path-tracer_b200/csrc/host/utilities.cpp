// utilities.cpp -- see utilities.h.  Behaviour follows src/utilities.cpp of the reference.
#include "utilities.h"
#include <sstream>

namespace utilityCore {

float clamp(float f, float lo, float hi) { return f < lo ? lo : (f > hi ? hi : f); }

std::string convertIntToString(int number) { return std::to_string(number); }

// whitespace-separated tokens (src/utilities.cpp:74-80)
std::vector<std::string> tokenizeString(std::string str) {
    std::vector<std::string> out;
    std::istringstream in(str);
    for (std::string tok; in >> tok;) out.push_back(tok);
    return out;
}

// transform = T * (Rx * Ry * Rz) * S with degrees turned into radians as rot * PI / 180 in float
// (src/utilities.cpp:65-72)
ptbglm::mat4 buildTransformationMatrix(ptbglm::vec3 translation, ptbglm::vec3 rotation, ptbglm::vec3 scale) {
    using namespace ptbglm;
    mat4 translationMat = translate(mat4(), translation);
    mat4 rotationMat = rotate(mat4(), rotation.x * (float)PI / 180, vec3(1, 0, 0));
    rotationMat = rotationMat * rotate(mat4(), rotation.y * (float)PI / 180, vec3(0, 1, 0));
    rotationMat = rotationMat * rotate(mat4(), rotation.z * (float)PI / 180, vec3(0, 0, 1));
    mat4 scaleMat = ptbglm::scale(mat4(), scale);
    return translationMat * rotationMat * scaleMat;
}

// line reader tolerant of \n, \r\n and \r endings; an empty read at end of input sets eofbit
// (src/utilities.cpp:82-112)
std::istream &safeGetline(std::istream &is, std::string &t) {
    t.clear();
    std::istream::sentry guard(is, true);
    std::streambuf *sb = is.rdbuf();
    while (true) {
        const int ch = sb->sbumpc();
        if (ch == '\n') break;
        if (ch == '\r') { if (sb->sgetc() == '\n') sb->sbumpc(); break; }
        if (ch == std::streambuf::traits_type::eof()) { if (t.empty()) is.setstate(std::ios::eofbit); break; }
        t.push_back((char)ch);
    }
    return is;
}

} // namespace utilityCore
