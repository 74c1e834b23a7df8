// utilities.h -- host utilities with the reference's names and behaviour (src/utilities.h, .cpp).
#pragma once
#include <istream>
#include <string>
#include <vector>
#include "glm_lite.h"

#define PI                3.1415926535897932384626422832795028841971f
#define TWO_PI            6.2831853071795864769252867665590057683943f
#define SQRT_OF_ONE_THIRD 0.5773502691896257645091487805019574556476f
#define EPSILON           0.00001f

namespace utilityCore {
    float clamp(float f, float min, float max);
    std::vector<std::string> tokenizeString(std::string str);
    ptbglm::mat4 buildTransformationMatrix(ptbglm::vec3 translation, ptbglm::vec3 rotation, ptbglm::vec3 scale);
    std::string convertIntToString(int number);
    std::istream &safeGetline(std::istream &is, std::string &t);
}
