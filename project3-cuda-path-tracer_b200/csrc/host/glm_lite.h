// glm_lite.h -- the handful of glm 0.9.6.3 types and functions the host side needs (scene loading,
// camera fix-up), restated with glm's exact operation order so that the matrices handed to the
// kernels are bit-identical to what the reference's loader produces.
// Citations are to external/include/glm/ in the reference.  Layouts match glm's (tightly packed
// floats, column-major mat4), so these structs can be memcpy'd to and from the reference's.
#pragma once
#include <cmath>

namespace ptbglm {

struct ivec2 { int x = 0, y = 0; };
struct vec2 { float x = 0, y = 0; vec2() {} vec2(float a, float b) : x(a), y(b) {} };

struct vec3 {
    float x = 0, y = 0, z = 0;                       // glm 0.9.6 zero-initialises (detail/type_vec3.inl:39-43)
    vec3() {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    // like glm's templated constructor: every argument is static_cast to float
    template <typename A, typename B, typename C> vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 &operator+=(vec3 &a, vec3 b) { a = a + b; return a; }

struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float &operator[](int i) { return (&x)[i]; }
    const float &operator[](int i) const { return (&x)[i]; }
};
inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator-(vec4 a, vec4 b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline vec4 operator*(vec4 a, vec4 b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator/(vec4 a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }

struct mat4 {
    vec4 c[4];                                       // identity by default (detail/type_mat4x4.inl:99-107)
    mat4() { c[0] = vec4(1, 0, 0, 0); c[1] = vec4(0, 1, 0, 0); c[2] = vec4(0, 0, 1, 0); c[3] = vec4(0, 0, 0, 1); }
    vec4 &operator[](int i) { return c[i]; }
    const vec4 &operator[](int i) const { return c[i]; }
};

// detail/func_geometric.inl:64-72, 134-141, 94-100, 154-159; func_exponential.inl:148-153
inline float dot(vec3 a, vec3 b) { vec3 t(a.x * b.x, a.y * b.y, a.z * b.z); return t.x + t.y + t.z; }
inline vec3 cross(vec3 x, vec3 y) { return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }

// detail/type_mat4x4.inl:686-706
inline mat4 operator*(const mat4 &a, const mat4 &b) {
    mat4 r;
    for (int j = 0; j < 4; ++j) r[j] = a[0] * b[j][0] + a[1] * b[j][1] + a[2] * b[j][2] + a[3] * b[j][3];
    return r;
}
// gtc/matrix_transform.inl:40-50
inline mat4 translate(const mat4 &m, vec3 v) {
    mat4 r = m;
    r[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
    return r;
}
// gtc/matrix_transform.inl:52-85
inline mat4 rotate(const mat4 &m, float angle, vec3 v) {
    const float a = angle, c = std::cos(a), s = std::sin(a);
    const vec3 axis = normalize(v);
    const vec3 temp = (1.0f - c) * axis;
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = 0 + temp[0] * axis[1] + s * axis[2];
    R[0][2] = 0 + temp[0] * axis[2] - s * axis[1];
    R[1][0] = 0 + temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = 0 + temp[1] * axis[2] + s * axis[0];
    R[2][0] = 0 + temp[2] * axis[0] + s * axis[1];
    R[2][1] = 0 + temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    mat4 r;
    r[0] = m[0] * R[0][0] + m[1] * R[0][1] + m[2] * R[0][2];
    r[1] = m[0] * R[1][0] + m[1] * R[1][1] + m[2] * R[1][2];
    r[2] = m[0] * R[2][0] + m[1] * R[2][1] + m[2] * R[2][2];
    r[3] = m[3];
    return r;
}
// gtc/matrix_transform.inl:122-134
inline mat4 scale(const mat4 &m, vec3 v) {
    mat4 r;
    r[0] = m[0] * v[0]; r[1] = m[1] * v[1]; r[2] = m[2] * v[2]; r[3] = m[3];
    return r;
}
// detail/type_mat4x4.inl:37-92: cofactor expansion, then multiply by 1/det
inline mat4 inverse(const mat4 &m) {
    float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    vec4 Fac0(Coef00, Coef00, Coef02, Coef03), Fac1(Coef04, Coef04, Coef06, Coef07), Fac2(Coef08, Coef08, Coef10, Coef11);
    vec4 Fac3(Coef12, Coef12, Coef14, Coef15), Fac4(Coef16, Coef16, Coef18, Coef19), Fac5(Coef20, Coef20, Coef22, Coef23);
    vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]), Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
    vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]), Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);
    vec4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2), Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
    vec4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5), Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);
    vec4 SignA(+1, -1, +1, -1), SignB(-1, +1, -1, +1);
    mat4 Inverse;
    Inverse[0] = Inv0 * SignA; Inverse[1] = Inv1 * SignB; Inverse[2] = Inv2 * SignA; Inverse[3] = Inv3 * SignB;
    vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
    vec4 Dot0(m[0] * Row0);
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float OneOverDeterminant = 1.0f / Dot1;
    mat4 r;
    for (int j = 0; j < 4; ++j) r[j] = Inverse[j] * OneOverDeterminant;
    return r;
}
// gtc/matrix_inverse.inl:95-156: its own cofactor code, then divide by det
inline mat4 inverseTranspose(const mat4 &m) {
    float S00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], S01 = m[2][1] * m[3][3] - m[3][1] * m[2][3], S02 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    float S03 = m[2][0] * m[3][3] - m[3][0] * m[2][3], S04 = m[2][0] * m[3][2] - m[3][0] * m[2][2], S05 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    float S06 = m[1][2] * m[3][3] - m[3][2] * m[1][3], S07 = m[1][1] * m[3][3] - m[3][1] * m[1][3], S08 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    float S09 = m[1][0] * m[3][3] - m[3][0] * m[1][3], S10 = m[1][0] * m[3][2] - m[3][0] * m[1][2], S11 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    float S12 = m[1][0] * m[3][1] - m[3][0] * m[1][1], S13 = m[1][2] * m[2][3] - m[2][2] * m[1][3], S14 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float S15 = m[1][1] * m[2][2] - m[2][1] * m[1][2], S16 = m[1][0] * m[2][3] - m[2][0] * m[1][3], S17 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float S18 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    mat4 I;
    I[0][0] = +(m[1][1] * S00 - m[1][2] * S01 + m[1][3] * S02);
    I[0][1] = -(m[1][0] * S00 - m[1][2] * S03 + m[1][3] * S04);
    I[0][2] = +(m[1][0] * S01 - m[1][1] * S03 + m[1][3] * S05);
    I[0][3] = -(m[1][0] * S02 - m[1][1] * S04 + m[1][2] * S05);
    I[1][0] = -(m[0][1] * S00 - m[0][2] * S01 + m[0][3] * S02);
    I[1][1] = +(m[0][0] * S00 - m[0][2] * S03 + m[0][3] * S04);
    I[1][2] = -(m[0][0] * S01 - m[0][1] * S03 + m[0][3] * S05);
    I[1][3] = +(m[0][0] * S02 - m[0][1] * S04 + m[0][2] * S05);
    I[2][0] = +(m[0][1] * S06 - m[0][2] * S07 + m[0][3] * S08);
    I[2][1] = -(m[0][0] * S06 - m[0][2] * S09 + m[0][3] * S10);
    I[2][2] = +(m[0][0] * S11 - m[0][1] * S09 + m[0][3] * S12);
    I[2][3] = -(m[0][0] * S08 - m[0][1] * S10 + m[0][2] * S12);
    I[3][0] = -(m[0][1] * S13 - m[0][2] * S14 + m[0][3] * S15);
    I[3][1] = +(m[0][0] * S13 - m[0][2] * S16 + m[0][3] * S17);
    I[3][2] = -(m[0][0] * S14 - m[0][1] * S16 + m[0][3] * S18);
    I[3][3] = +(m[0][0] * S15 - m[0][1] * S17 + m[0][2] * S18);
    float Determinant = +m[0][0] * I[0][0] + m[0][1] * I[0][1] + m[0][2] * I[0][2] + m[0][3] * I[0][3];
    for (int j = 0; j < 4; ++j) I[j] = I[j] / Determinant;
    return I;
}

} // namespace ptbglm
