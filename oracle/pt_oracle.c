/*
 * pt_oracle.c -- TEST INFRASTRUCTURE (kind = "port").  Not part of the product.
 *
 * A plain-C11 restatement of the reference's path-tracing hot path with no glm, no thrust and no
 * reference headers, so that it builds and runs on machines where /root/reference is absent (the
 * GPU box).  Every function cites the reference lines it follows; the glm 0.9.6.3 operation order
 * is spelled out term by term because bit-exact first-bounce parity depends on it (SURVEY.md 7,
 * hard part 1).  Build: oracle/Makefile, pinned to -O2 -ffp-contract=off (no FMA contraction).
 *
 * PINNED: tests/test_oracle.py checks this file against (a) the known answers the reference-linked
 * build (oracle/_ref/libptref.so = the reference's own intersections.h / interactions.h / scene.cpp
 * compiled here) produced -- SURVEY.md 8(c)/9 -- and (b) bit for bit against that library itself
 * whenever it is present, and (c) against the committed fixtures under tests/golden/ that
 * tests/golden/make_golden.py generated from the reference-linked build.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker.
 */
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "orc_api.h"

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;
typedef struct { v4 c[4]; } m4;      /* column-major like glm: c[col].row */

/* ------------------------------------------------------------------ glm vector restatements */
static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v4 V4(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }
static inline v3 v3add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3sub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3mul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3scale(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 v3neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline v4 v4add(v4 a, v4 b) { return V4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline v4 v4sub(v4 a, v4 b) { return V4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline v4 v4mul(v4 a, v4 b) { return V4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline v4 v4scale(v4 a, float s) { return V4(a.x * s, a.y * s, a.z * s, a.w * s); }

/* glm/detail/func_geometric.inl:64-72  dot(vec3) = (x*x' + y*y') + z*z' */
static inline float v3dot(v3 a, v3 b) {
    float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return tx + ty + tz;
}
/* func_geometric.inl:134-141 */
static inline v3 v3cross(v3 x, v3 y) {
    return V3(x.y * y.z - y.y * x.z,
              x.z * y.x - y.z * x.x,
              x.x * y.y - y.x * x.y);
}
/* func_geometric.inl:154-159 with func_exponential.inl:148-153: v * (1 / sqrt(dot(v,v))) */
static inline v3 v3normalize(v3 v) {
    float inv = 1.0f / sqrtf(v3dot(v, v));
    return V3(v.x * inv, v.y * inv, v.z * inv);
}
/* func_geometric.inl:94-100 */
static inline float v3length(v3 v) { return sqrtf(v3dot(v, v)); }

/* glm/detail/type_mat4x4.inl:617-628  mat4 * vec4 = (m0*v0 + m1*v1) + (m2*v2 + m3*v3), clipped to
 * vec3 as multiplyMV does (src/intersections.h:34-36) */
static inline v3 mulMV(const float *m, v3 v, float w) {
    v3 r;
    r.x = (m[0] * v.x + m[4] * v.y) + (m[8]  * v.z + m[12] * w);
    r.y = (m[1] * v.x + m[5] * v.y) + (m[9]  * v.z + m[13] * w);
    r.z = (m[2] * v.x + m[6] * v.y) + (m[10] * v.z + m[14] * w);
    return r;
}

/* ------------------------------------------------------------------ glm matrix restatements */
static m4 m4identity(void) {
    m4 r;
    r.c[0] = V4(1, 0, 0, 0); r.c[1] = V4(0, 1, 0, 0); r.c[2] = V4(0, 0, 1, 0); r.c[3] = V4(0, 0, 0, 1);
    return r;
}
/* type_mat4x4.inl:686-706: Result[j] = ((A0*b[j][0] + A1*b[j][1]) + A2*b[j][2]) + A3*b[j][3] */
static m4 m4mul(m4 a, m4 b) {
    m4 r;
    for (int j = 0; j < 4; j++) {
        v4 t = v4add(v4scale(a.c[0], b.c[j].x), v4scale(a.c[1], b.c[j].y));
        t = v4add(t, v4scale(a.c[2], b.c[j].z));
        t = v4add(t, v4scale(a.c[3], b.c[j].w));
        r.c[j] = t;
    }
    return r;
}
/* gtc/matrix_transform.inl:40-50 */
static m4 m4translate(m4 m, v3 v) {
    m4 r = m;
    v4 t = v4add(v4scale(m.c[0], v.x), v4scale(m.c[1], v.y));
    t = v4add(t, v4scale(m.c[2], v.z));
    r.c[3] = v4add(t, m.c[3]);
    return r;
}
/* gtc/matrix_transform.inl:52-85 */
static m4 m4rotate(m4 m, float angle, v3 v) {
    const float a = angle;
    const float c = cosf(a);
    const float s = sinf(a);
    v3 axis = v3normalize(v);
    v3 temp = V3((1.0f - c) * axis.x, (1.0f - c) * axis.y, (1.0f - c) * axis.z);
    float R00 = c + temp.x * axis.x;
    float R01 = 0.0f + temp.x * axis.y + s * axis.z;
    float R02 = 0.0f + temp.x * axis.z - s * axis.y;
    float R10 = 0.0f + temp.y * axis.x - s * axis.z;
    float R11 = c + temp.y * axis.y;
    float R12 = 0.0f + temp.y * axis.z + s * axis.x;
    float R20 = 0.0f + temp.z * axis.x + s * axis.y;
    float R21 = 0.0f + temp.z * axis.y - s * axis.x;
    float R22 = c + temp.z * axis.z;
    m4 r;
    r.c[0] = v4add(v4add(v4scale(m.c[0], R00), v4scale(m.c[1], R01)), v4scale(m.c[2], R02));
    r.c[1] = v4add(v4add(v4scale(m.c[0], R10), v4scale(m.c[1], R11)), v4scale(m.c[2], R12));
    r.c[2] = v4add(v4add(v4scale(m.c[0], R20), v4scale(m.c[1], R21)), v4scale(m.c[2], R22));
    r.c[3] = m.c[3];
    return r;
}
/* gtc/matrix_transform.inl:122-134 */
static m4 m4scale(m4 m, v3 v) {
    m4 r;
    r.c[0] = v4scale(m.c[0], v.x);
    r.c[1] = v4scale(m.c[1], v.y);
    r.c[2] = v4scale(m.c[2], v.z);
    r.c[3] = m.c[3];
    return r;
}
#define M(mat, col, row) (((const float *)&(mat).c[col])[row])
/* type_mat4x4.inl:37-92  (cofactors, then multiply by 1/det) */
static m4 m4inverse(m4 m) {
    float Coef00 = M(m,2,2) * M(m,3,3) - M(m,3,2) * M(m,2,3);
    float Coef02 = M(m,1,2) * M(m,3,3) - M(m,3,2) * M(m,1,3);
    float Coef03 = M(m,1,2) * M(m,2,3) - M(m,2,2) * M(m,1,3);
    float Coef04 = M(m,2,1) * M(m,3,3) - M(m,3,1) * M(m,2,3);
    float Coef06 = M(m,1,1) * M(m,3,3) - M(m,3,1) * M(m,1,3);
    float Coef07 = M(m,1,1) * M(m,2,3) - M(m,2,1) * M(m,1,3);
    float Coef08 = M(m,2,1) * M(m,3,2) - M(m,3,1) * M(m,2,2);
    float Coef10 = M(m,1,1) * M(m,3,2) - M(m,3,1) * M(m,1,2);
    float Coef11 = M(m,1,1) * M(m,2,2) - M(m,2,1) * M(m,1,2);
    float Coef12 = M(m,2,0) * M(m,3,3) - M(m,3,0) * M(m,2,3);
    float Coef14 = M(m,1,0) * M(m,3,3) - M(m,3,0) * M(m,1,3);
    float Coef15 = M(m,1,0) * M(m,2,3) - M(m,2,0) * M(m,1,3);
    float Coef16 = M(m,2,0) * M(m,3,2) - M(m,3,0) * M(m,2,2);
    float Coef18 = M(m,1,0) * M(m,3,2) - M(m,3,0) * M(m,1,2);
    float Coef19 = M(m,1,0) * M(m,2,2) - M(m,2,0) * M(m,1,2);
    float Coef20 = M(m,2,0) * M(m,3,1) - M(m,3,0) * M(m,2,1);
    float Coef22 = M(m,1,0) * M(m,3,1) - M(m,3,0) * M(m,1,1);
    float Coef23 = M(m,1,0) * M(m,2,1) - M(m,2,0) * M(m,1,1);

    v4 Fac0 = V4(Coef00, Coef00, Coef02, Coef03);
    v4 Fac1 = V4(Coef04, Coef04, Coef06, Coef07);
    v4 Fac2 = V4(Coef08, Coef08, Coef10, Coef11);
    v4 Fac3 = V4(Coef12, Coef12, Coef14, Coef15);
    v4 Fac4 = V4(Coef16, Coef16, Coef18, Coef19);
    v4 Fac5 = V4(Coef20, Coef20, Coef22, Coef23);

    v4 Vec0 = V4(M(m,1,0), M(m,0,0), M(m,0,0), M(m,0,0));
    v4 Vec1 = V4(M(m,1,1), M(m,0,1), M(m,0,1), M(m,0,1));
    v4 Vec2 = V4(M(m,1,2), M(m,0,2), M(m,0,2), M(m,0,2));
    v4 Vec3 = V4(M(m,1,3), M(m,0,3), M(m,0,3), M(m,0,3));

    v4 Inv0 = v4add(v4sub(v4mul(Vec1, Fac0), v4mul(Vec2, Fac1)), v4mul(Vec3, Fac2));
    v4 Inv1 = v4add(v4sub(v4mul(Vec0, Fac0), v4mul(Vec2, Fac3)), v4mul(Vec3, Fac4));
    v4 Inv2 = v4add(v4sub(v4mul(Vec0, Fac1), v4mul(Vec1, Fac3)), v4mul(Vec3, Fac5));
    v4 Inv3 = v4add(v4sub(v4mul(Vec0, Fac2), v4mul(Vec1, Fac4)), v4mul(Vec2, Fac5));

    v4 SignA = V4(+1, -1, +1, -1);
    v4 SignB = V4(-1, +1, -1, +1);
    m4 Inverse;
    Inverse.c[0] = v4mul(Inv0, SignA);
    Inverse.c[1] = v4mul(Inv1, SignB);
    Inverse.c[2] = v4mul(Inv2, SignA);
    Inverse.c[3] = v4mul(Inv3, SignB);

    v4 Row0 = V4(Inverse.c[0].x, Inverse.c[1].x, Inverse.c[2].x, Inverse.c[3].x);
    v4 Dot0 = v4mul(m.c[0], Row0);
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float OneOverDeterminant = 1.0f / Dot1;
    m4 r;
    for (int j = 0; j < 4; j++) r.c[j] = v4scale(Inverse.c[j], OneOverDeterminant);
    return r;
}
/* gtc/matrix_inverse.inl:95-156  (separate cofactor code, then divide by det) */
static m4 m4inverseTranspose(m4 m) {
    float S00 = M(m,2,2) * M(m,3,3) - M(m,3,2) * M(m,2,3);
    float S01 = M(m,2,1) * M(m,3,3) - M(m,3,1) * M(m,2,3);
    float S02 = M(m,2,1) * M(m,3,2) - M(m,3,1) * M(m,2,2);
    float S03 = M(m,2,0) * M(m,3,3) - M(m,3,0) * M(m,2,3);
    float S04 = M(m,2,0) * M(m,3,2) - M(m,3,0) * M(m,2,2);
    float S05 = M(m,2,0) * M(m,3,1) - M(m,3,0) * M(m,2,1);
    float S06 = M(m,1,2) * M(m,3,3) - M(m,3,2) * M(m,1,3);
    float S07 = M(m,1,1) * M(m,3,3) - M(m,3,1) * M(m,1,3);
    float S08 = M(m,1,1) * M(m,3,2) - M(m,3,1) * M(m,1,2);
    float S09 = M(m,1,0) * M(m,3,3) - M(m,3,0) * M(m,1,3);
    float S10 = M(m,1,0) * M(m,3,2) - M(m,3,0) * M(m,1,2);
    float S11 = M(m,1,1) * M(m,3,3) - M(m,3,1) * M(m,1,3);
    float S12 = M(m,1,0) * M(m,3,1) - M(m,3,0) * M(m,1,1);
    float S13 = M(m,1,2) * M(m,2,3) - M(m,2,2) * M(m,1,3);
    float S14 = M(m,1,1) * M(m,2,3) - M(m,2,1) * M(m,1,3);
    float S15 = M(m,1,1) * M(m,2,2) - M(m,2,1) * M(m,1,2);
    float S16 = M(m,1,0) * M(m,2,3) - M(m,2,0) * M(m,1,3);
    float S17 = M(m,1,0) * M(m,2,2) - M(m,2,0) * M(m,1,2);
    float S18 = M(m,1,0) * M(m,2,1) - M(m,2,0) * M(m,1,1);

    float I[4][4];
    I[0][0] = +(M(m,1,1) * S00 - M(m,1,2) * S01 + M(m,1,3) * S02);
    I[0][1] = -(M(m,1,0) * S00 - M(m,1,2) * S03 + M(m,1,3) * S04);
    I[0][2] = +(M(m,1,0) * S01 - M(m,1,1) * S03 + M(m,1,3) * S05);
    I[0][3] = -(M(m,1,0) * S02 - M(m,1,1) * S04 + M(m,1,2) * S05);

    I[1][0] = -(M(m,0,1) * S00 - M(m,0,2) * S01 + M(m,0,3) * S02);
    I[1][1] = +(M(m,0,0) * S00 - M(m,0,2) * S03 + M(m,0,3) * S04);
    I[1][2] = -(M(m,0,0) * S01 - M(m,0,1) * S03 + M(m,0,3) * S05);
    I[1][3] = +(M(m,0,0) * S02 - M(m,0,1) * S04 + M(m,0,2) * S05);

    I[2][0] = +(M(m,0,1) * S06 - M(m,0,2) * S07 + M(m,0,3) * S08);
    I[2][1] = -(M(m,0,0) * S06 - M(m,0,2) * S09 + M(m,0,3) * S10);
    I[2][2] = +(M(m,0,0) * S11 - M(m,0,1) * S09 + M(m,0,3) * S12);
    I[2][3] = -(M(m,0,0) * S08 - M(m,0,1) * S10 + M(m,0,2) * S12);

    I[3][0] = -(M(m,0,1) * S13 - M(m,0,2) * S14 + M(m,0,3) * S15);
    I[3][1] = +(M(m,0,0) * S13 - M(m,0,2) * S16 + M(m,0,3) * S17);
    I[3][2] = -(M(m,0,0) * S14 - M(m,0,1) * S16 + M(m,0,3) * S18);
    I[3][3] = +(M(m,0,0) * S15 - M(m,0,1) * S17 + M(m,0,2) * S18);

    float Determinant = +M(m,0,0) * I[0][0] + M(m,0,1) * I[0][1] + M(m,0,2) * I[0][2] + M(m,0,3) * I[0][3];
    m4 r;
    for (int c = 0; c < 4; c++) {
        r.c[c] = V4(I[c][0] / Determinant, I[c][1] / Determinant, I[c][2] / Determinant, I[c][3] / Determinant);
    }
    return r;
}

#define PI_F      3.1415926535897932384626422832795028841971f   /* src/utilities.h:12 */
#define TWO_PI_F  6.2831853071795864769252867665590057683943f   /* src/utilities.h:13 */
#define SQRT_OF_ONE_THIRD_F 0.5773502691896257645091487805019574556476f /* src/utilities.h:14 */

/* src/utilities.cpp:65-72: T * (Rx * Ry * Rz) * S, degrees -> radians as rot * PI / 180 in float */
static m4 buildTransformationMatrix(v3 translation, v3 rotation, v3 scale) {
    m4 I = m4identity();
    m4 translationMat = m4translate(I, translation);
    m4 rotationMat = m4rotate(I, rotation.x * PI_F / 180, V3(1, 0, 0));
    rotationMat = m4mul(rotationMat, m4rotate(I, rotation.y * PI_F / 180, V3(0, 1, 0)));
    rotationMat = m4mul(rotationMat, m4rotate(I, rotation.z * PI_F / 180, V3(0, 0, 1)));
    m4 scaleMat = m4scale(I, scale);
    return m4mul(m4mul(translationMat, rotationMat), scaleMat);
}

/* ------------------------------------------------------------------ scene file (src/scene.cpp) */
typedef struct {
    OrcGeom *geoms; int ngeoms, capGeoms;
    OrcMaterial *mats; int nmats, capMats;
    OrcCamera cam;
    int iterations, traceDepth;
    char imageName[256];
} OScene;

typedef struct { const char *buf; size_t len, pos; int good; } LineReader;

/* src/utilities.cpp:82-112 safeGetline: accepts \n, \r\n and \r; sets "not good" only when an
 * empty read hits EOF */
static void readLine(LineReader *lr, char *out, size_t cap) {
    size_t n = 0;
    out[0] = 0;
    for (;;) {
        if (lr->pos >= lr->len) {
            if (n == 0) lr->good = 0;
            break;
        }
        char c = lr->buf[lr->pos++];
        if (c == '\n') break;
        if (c == '\r') {
            if (lr->pos < lr->len && lr->buf[lr->pos] == '\n') lr->pos++;
            break;
        }
        if (n + 1 < cap) out[n++] = c;
    }
    out[n] = 0;
}

/* src/utilities.cpp:74-80 tokenizeString: split on whitespace */
static int tokenize(char *line, char **tok, int maxTok) {
    int n = 0;
    char *p = line;
    while (*p) {
        while (*p == ' ' || *p == '\t' || *p == '\v' || *p == '\f' || *p == '\r' || *p == '\n') p++;
        if (!*p) break;
        if (n < maxTok) tok[n++] = p;
        while (*p && !(*p == ' ' || *p == '\t' || *p == '\v' || *p == '\f' || *p == '\r' || *p == '\n')) p++;
        if (*p) *p++ = 0;
    }
    return n;
}
static float tokf(char **tok, int n, int i) { return i < n ? (float)atof(tok[i]) : 0.0f; }

/* src/scene.cpp:153-188 */
static void loadMaterial(OScene *s, LineReader *lr, const char *idTok) {
    int id = atoi(idTok);
    if (id != s->nmats) return;                       /* :155-157 */
    OrcMaterial m;
    memset(&m, 0, sizeof m);
    char line[1024], *tok[16];
    for (int i = 0; i < 7; i++) {                     /* exactly seven property lines, :163 */
        readLine(lr, line, sizeof line);
        int n = tokenize(line, tok, 16);
        if (n == 0) continue;
        if (!strcmp(tok[0], "RGB")) { m.color[0] = tokf(tok, n, 1); m.color[1] = tokf(tok, n, 2); m.color[2] = tokf(tok, n, 3); }
        else if (!strcmp(tok[0], "SPECEX")) m.specExponent = tokf(tok, n, 1);
        else if (!strcmp(tok[0], "SPECRGB")) { m.specColor[0] = tokf(tok, n, 1); m.specColor[1] = tokf(tok, n, 2); m.specColor[2] = tokf(tok, n, 3); }
        else if (!strcmp(tok[0], "REFL")) m.hasReflective = tokf(tok, n, 1);
        else if (!strcmp(tok[0], "REFR")) m.hasRefractive = tokf(tok, n, 1);
        else if (!strcmp(tok[0], "REFRIOR")) m.indexOfRefraction = tokf(tok, n, 1);
        else if (!strcmp(tok[0], "EMITTANCE")) m.emittance = tokf(tok, n, 1);
    }
    if (s->nmats == s->capMats) {
        s->capMats = s->capMats ? 2 * s->capMats : 16;
        s->mats = (OrcMaterial *)realloc(s->mats, (size_t)s->capMats * sizeof(OrcMaterial));
    }
    s->mats[s->nmats++] = m;
}

static void m4store(float *dst, m4 m) { memcpy(dst, &m, 64); }

/* src/scene.cpp:35-90 */
static void loadGeom(OScene *s, LineReader *lr, const char *idTok) {
    int id = atoi(idTok);
    if (id != s->ngeoms) return;                      /* :37-39 */
    OrcGeom g;
    memset(&g, 0, sizeof g);
    char line[1024], *tok[16];
    readLine(lr, line, sizeof line);                  /* type: compared as a whole line, :46-55 */
    if (line[0] && lr->good) {
        if (!strcmp(line, "sphere")) g.type = 0;
        else if (!strcmp(line, "cube")) g.type = 1;
    }
    readLine(lr, line, sizeof line);                  /* "material N", :58-63 */
    if (line[0] && lr->good) {
        int n = tokenize(line, tok, 16);
        if (n > 1) g.materialid = atoi(tok[1]);
    }
    readLine(lr, line, sizeof line);                  /* TRANS / ROTAT / SCALE until a blank line, :66-80 */
    while (line[0] && lr->good) {
        int n = tokenize(line, tok, 16);
        if (n > 0) {
            float *dst = NULL;
            if (!strcmp(tok[0], "TRANS")) dst = g.translation;
            else if (!strcmp(tok[0], "ROTAT")) dst = g.rotation;
            else if (!strcmp(tok[0], "SCALE")) dst = g.scale;
            if (dst) { dst[0] = tokf(tok, n, 1); dst[1] = tokf(tok, n, 2); dst[2] = tokf(tok, n, 3); }
        }
        readLine(lr, line, sizeof line);
    }
    m4 T = buildTransformationMatrix(V3(g.translation[0], g.translation[1], g.translation[2]),
                                     V3(g.rotation[0], g.rotation[1], g.rotation[2]),
                                     V3(g.scale[0], g.scale[1], g.scale[2]));   /* :82-83 */
    m4store(g.transform, T);
    m4store(g.inverseTransform, m4inverse(T));                                   /* :84 */
    m4store(g.invTranspose, m4inverseTranspose(T));                              /* :85 */
    if (s->ngeoms == s->capGeoms) {
        s->capGeoms = s->capGeoms ? 2 * s->capGeoms : 64;
        s->geoms = (OrcGeom *)realloc(s->geoms, (size_t)s->capGeoms * sizeof(OrcGeom));
    }
    s->geoms[s->ngeoms++] = g;
}

/* src/scene.cpp:92-151 */
static void loadCamera(OScene *s, LineReader *lr) {
    OrcCamera *cam = &s->cam;
    float fovy = 0.0f;
    char line[1024], *tok[16];
    for (int i = 0; i < 5; i++) {                     /* five fixed lines, :99 */
        readLine(lr, line, sizeof line);
        int n = tokenize(line, tok, 16);
        if (n == 0) continue;
        if (!strcmp(tok[0], "RES")) { cam->resolution[0] = n > 1 ? atoi(tok[1]) : 0; cam->resolution[1] = n > 2 ? atoi(tok[2]) : 0; }
        else if (!strcmp(tok[0], "FOVY")) fovy = tokf(tok, n, 1);
        else if (!strcmp(tok[0], "ITERATIONS")) s->iterations = n > 1 ? atoi(tok[1]) : 0;
        else if (!strcmp(tok[0], "DEPTH")) s->traceDepth = n > 1 ? atoi(tok[1]) : 0;
        else if (!strcmp(tok[0], "FILE")) { if (n > 1) { strncpy(s->imageName, tok[1], sizeof s->imageName - 1); } }
    }
    readLine(lr, line, sizeof line);
    while (line[0] && lr->good) {                     /* EYE / LOOKAT / UP until a blank line, :117-130 */
        int n = tokenize(line, tok, 16);
        if (n > 0) {
            float *dst = NULL;
            if (!strcmp(tok[0], "EYE")) dst = cam->position;
            else if (!strcmp(tok[0], "LOOKAT")) dst = cam->lookAt;
            else if (!strcmp(tok[0], "UP")) dst = cam->up;
            if (dst) { dst[0] = tokf(tok, n, 1); dst[1] = tokf(tok, n, 2); dst[2] = tokf(tok, n, 3); }
        }
        readLine(lr, line, sizeof line);
    }
    /* :133-140 */
    float yscaled = tanf(fovy * (PI_F / 180));
    float xscaled = (yscaled * cam->resolution[0]) / cam->resolution[1];
    float fovx = (atanf(xscaled) * 180) / PI_F;
    cam->fov[0] = fovx; cam->fov[1] = fovy;
    /* :138 -- right is computed from `view` BEFORE view is assigned (view is still zero => NaN) */
    v3 right = v3normalize(v3cross(V3(cam->view[0], cam->view[1], cam->view[2]), V3(cam->up[0], cam->up[1], cam->up[2])));
    cam->right[0] = right.x; cam->right[1] = right.y; cam->right[2] = right.z;
    cam->pixelLength[0] = 2 * xscaled / (float)cam->resolution[0];
    cam->pixelLength[1] = 2 * yscaled / (float)cam->resolution[1];
    /* :142 */
    v3 view = v3normalize(v3sub(V3(cam->lookAt[0], cam->lookAt[1], cam->lookAt[2]),
                                V3(cam->position[0], cam->position[1], cam->position[2])));
    cam->view[0] = view.x; cam->view[1] = view.y; cam->view[2] = view.z;
}

const char *orc_kind(void) { return "port"; }

/* src/scene.cpp:7-33 */
void *orc_scene_load(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)len + 1);
    if (fread(buf, 1, (size_t)len, f) != (size_t)len) { fclose(f); free(buf); return NULL; }
    fclose(f);
    OScene *s = (OScene *)calloc(1, sizeof(OScene));
    LineReader lr = { buf, (size_t)len, 0, 1 };
    char line[1024], *tok[16];
    while (lr.good) {
        readLine(&lr, line, sizeof line);
        if (line[0]) {
            int n = tokenize(line, tok, 16);
            if (n == 0) continue;
            if (!strcmp(tok[0], "MATERIAL") && n > 1) loadMaterial(s, &lr, tok[1]);
            else if (!strcmp(tok[0], "OBJECT") && n > 1) loadGeom(s, &lr, tok[1]);
            else if (!strcmp(tok[0], "CAMERA")) loadCamera(s, &lr);
        }
    }
    free(buf);
    return s;
}

void orc_scene_counts(void *scene, int *ngeoms, int *nmaterials) {
    OScene *s = (OScene *)scene;
    *ngeoms = s->ngeoms; *nmaterials = s->nmats;
}

void orc_scene_get(void *scene, OrcGeom *geoms, OrcMaterial *materials, OrcCamera *cam,
                   int *iterations, int *traceDepth, char *imageName, int nameCap) {
    OScene *s = (OScene *)scene;
    if (geoms) memcpy(geoms, s->geoms, (size_t)s->ngeoms * sizeof(OrcGeom));
    if (materials) memcpy(materials, s->mats, (size_t)s->nmats * sizeof(OrcMaterial));
    if (cam) *cam = s->cam;
    if (iterations) *iterations = s->iterations;
    if (traceDepth) *traceDepth = s->traceDepth;
    if (imageName && nameCap > 0) { strncpy(imageName, s->imageName, (size_t)nameCap - 1); imageName[nameCap - 1] = 0; }
}

void orc_scene_free(void *scene) {
    OScene *s = (OScene *)scene;
    if (!s) return;
    free(s->geoms); free(s->mats); free(s);
}

/* src/main.cpp:53-67 then :102-120 */
void orc_camera_fixup(OrcCamera *cam) {
    v3 view = V3(cam->view[0], cam->view[1], cam->view[2]);
    v3 lookAt = V3(cam->lookAt[0], cam->lookAt[1], cam->lookAt[2]);
    v3 position = V3(cam->position[0], cam->position[1], cam->position[2]);
    v3 viewXZ = V3(view.x, 0.0f, view.z);
    v3 viewZY = V3(0.0f, view.y, view.z);
    float phi = acosf(v3dot(v3normalize(viewXZ), V3(0, 0, -1)));      /* :64 */
    float theta = acosf(v3dot(v3normalize(viewZY), V3(0, 1, 0)));     /* :65 */
    float zoom = v3length(v3sub(position, lookAt));                   /* :67 */

    v3 cp;
    cp.x = zoom * sinf(phi) * sinf(theta);                            /* :105-107 */
    cp.y = zoom * cosf(theta);
    cp.z = zoom * cosf(phi) * sinf(theta);

    v3 v = v3neg(v3normalize(cp));                                    /* :109 */
    v3 u = V3(0, 1, 0);
    v3 r = v3cross(v, u);                                             /* :112 (not normalised) */
    v3 up = v3cross(r, v);                                            /* :113 */
    cp = v3add(cp, lookAt);                                           /* :117 */
    cam->view[0] = v.x; cam->view[1] = v.y; cam->view[2] = v.z;
    cam->up[0] = up.x; cam->up[1] = up.y; cam->up[2] = up.z;
    cam->right[0] = r.x; cam->right[1] = r.y; cam->right[2] = r.z;
    cam->position[0] = cp.x; cam->position[1] = cp.y; cam->position[2] = cp.z;
}

void orc_build_transform(const float t[3], const float r[3], const float s[3],
                         float transform[16], float inverse[16], float invTranspose[16]) {
    m4 T = buildTransformationMatrix(V3(t[0], t[1], t[2]), V3(r[0], r[1], r[2]), V3(s[0], s[1], s[2]));
    m4store(transform, T);
    m4store(inverse, m4inverse(T));
    m4store(invTranspose, m4inverseTranspose(T));
}

/* ------------------------------------------------------------------ RNG
 * src/intersections.h:12-20 utilhash; src/pathtrace.cu:41-45 makeSeededRandomEngine;
 * thrust::default_random_engine = minstd_rand = LCG(a = 48271, c = 0, m = 2^31 - 1)
 * (thrust/random/linear_congruential_engine.h, CUDA 12.9 / Thrust 2.8.2; the reference asks for
 * CUDA >= 8.0, CMakeLists.txt:65).  Published algorithm restated:
 *   seed(s):  x = s mod m, and 1 when that is 0              (detail/linear_congruential_engine.inl)
 *   next():   x = a*x mod m (Schrage in 32 bits == exact product mod m)   (detail/mod.h)
 *   u01:      float(x - 1) / (1.0f + float(m - 2)) = float(x-1) / 2147483648.0f, then *(b-a)+a
 *                                                            (detail/uniform_real_distribution.inl) */
static inline unsigned int utilhash(unsigned int a) {
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    return a;
}
typedef struct { unsigned int x; } Rng;
static inline Rng seededEngine(int iter, int index, int depth) {
    unsigned int h = utilhash((1u << 31) | ((unsigned int)depth << 22) | (unsigned int)iter) ^ utilhash((unsigned int)index);
    Rng r;
    r.x = h % 2147483647u;
    if (r.x == 0) r.x = 1;
    return r;
}
static inline float u01(Rng *r) {
    r->x = (unsigned int)(((unsigned long long)r->x * 48271ull) % 2147483647ull);
    float result = (float)(r->x - 1u);
    result /= (1.0f + (float)(2147483646u - 1u));
    return (result * (1.0f - 0.0f)) + 0.0f;
}

void orc_rng_u01(int iter, int index, int depth, int n, float *out) {
    Rng r = seededEngine(iter, index, depth);
    for (int i = 0; i < n; i++) out[i] = u01(&r);
}

/* ------------------------------------------------------------------ intersections (src/intersections.h) */
typedef struct { v3 o, d; } ORay;

/* :27-29 */
static inline v3 getPointOnRay(ORay r, float t) {
    return v3add(r.o, v3scale(v3normalize(r.d), t - .0001f));
}

/* :48-90 */
static float boxIntersectionTest(const OrcGeom *box, ORay r, v3 *normal, int *outside) {
    ORay q;
    q.o = mulMV(box->inverseTransform, r.o, 1.0f);
    q.d = v3normalize(mulMV(box->inverseTransform, r.d, 0.0f));

    float tmin = -1e38f, tmax = 1e38f;
    float tmin_n[3] = {0, 0, 0}, tmax_n[3] = {0, 0, 0};
    const float qo[3] = {q.o.x, q.o.y, q.o.z}, qd[3] = {q.d.x, q.d.y, q.d.z};
    for (int xyz = 0; xyz < 3; ++xyz) {
        float qdxyz = qd[xyz];
        float t1 = (-0.5f - qo[xyz]) / qdxyz;          /* unguarded division, :60-62 */
        float t2 = (+0.5f - qo[xyz]) / qdxyz;
        float ta = t1 < t2 ? t1 : t2;                  /* glm::min(x,y) = x<y?x:y (func_common.inl:409-414) */
        float tb = t1 > t2 ? t1 : t2;                  /* glm::max(x,y) = x>y?x:y (:430-435) */
        float n[3] = {0, 0, 0};
        n[xyz] = t2 < t1 ? +1.0f : -1.0f;
        if (ta > 0 && ta > tmin) { tmin = ta; memcpy(tmin_n, n, sizeof n); }
        if (tb < tmax) { tmax = tb; memcpy(tmax_n, n, sizeof n); }
    }
    if (tmax >= tmin && tmax > 0) {
        *outside = 1;
        if (tmin <= 0) {
            tmin = tmax;
            memcpy(tmin_n, tmax_n, sizeof tmin_n);
            *outside = 0;
        }
        v3 P = mulMV(box->transform, getPointOnRay(q, tmin), 1.0f);
        *normal = v3normalize(mulMV(box->transform, V3(tmin_n[0], tmin_n[1], tmin_n[2]), 0.0f));  /* :86 uses transform */
        return v3length(v3sub(r.o, P));
    }
    return -1;
}

/* :102-144 */
static float sphereIntersectionTest(const OrcGeom *sphere, ORay r, v3 *normal, int *outside) {
    const float radius = .5f;
    ORay rt;
    rt.o = mulMV(sphere->inverseTransform, r.o, 1.0f);
    rt.d = v3normalize(mulMV(sphere->inverseTransform, r.d, 0.0f));

    float vDotDirection = v3dot(rt.o, rt.d);
    float radicand = vDotDirection * vDotDirection - (v3dot(rt.o, rt.o) - powf(radius, 2));
    if (radicand < 0) return -1;

    float squareRoot = sqrtf(radicand);
    float firstTerm = -vDotDirection;
    float t1 = firstTerm + squareRoot;
    float t2 = firstTerm - squareRoot;

    float t = 0;
    if (t1 < 0 && t2 < 0) {
        return -1;
    } else if (t1 > 0 && t2 > 0) {
        t = t2 < t1 ? t2 : t1;                         /* std::min on the host build, :128 */
        *outside = 1;
    } else {
        t = t1 < t2 ? t2 : t1;                         /* std::max on the host build, :131 */
        *outside = 0;
    }
    v3 objP = getPointOnRay(rt, t);
    v3 P = mulMV(sphere->transform, objP, 1.f);
    v3 n = v3normalize(mulMV(sphere->invTranspose, objP, 0.f));
    if (!*outside) n = v3neg(n);
    *normal = n;
    return v3length(v3sub(r.o, P));
}

/* glm::intersectRayTriangle, glm/gtx/intersect.inl:29-66 of the vendored glm 0.9.6.3 (Moeller-Trumbore,
 * back faces culled by `a < epsilon`); bary = (u, v, t) */
static int intersectRayTriangle(v3 orig, v3 dir, v3 v0, v3 v1, v3 v2, v3 *bary) {
    v3 e1 = v3sub(v1, v0);
    v3 e2 = v3sub(v2, v0);
    v3 p = v3cross(dir, e2);
    float a = v3dot(e1, p);
    if (a < FLT_EPSILON) return 0;
    float f = 1.0f / a;
    v3 s = v3sub(orig, v0);
    bary->x = f * v3dot(s, p);
    if (bary->x < 0.0f) return 0;
    if (bary->x > 1.0f) return 0;
    v3 q = v3cross(s, e1);
    bary->y = f * v3dot(dir, q);
    if (bary->y < 0.0f) return 0;
    if (bary->y + bary->x > 1.0f) return 0;
    bary->z = f * v3dot(e2, q);
    return bary->z >= 0.0f;
}

/* The mesh test (authored, see orc_api.h), built like boxIntersectionTest / sphereIntersectionTest. */
static float meshIntersectionTest(const OrcGeom *mesh, const OrcTriangle *tris, int count, ORay r, v3 *normal, int *outside) {
    ORay q;
    q.o = mulMV(mesh->inverseTransform, r.o, 1.0f);
    q.d = v3normalize(mulMV(mesh->inverseTransform, r.d, 0.0f));
    float tmin = FLT_MAX;
    int best = -1, bestFront = 1;
    for (int i = 0; i < count; i++) {
        v3 v0 = V3(tris[i].v0[0], tris[i].v0[1], tris[i].v0[2]);
        v3 v1 = V3(tris[i].v1[0], tris[i].v1[1], tris[i].v1[2]);
        v3 v2 = V3(tris[i].v2[0], tris[i].v2[1], tris[i].v2[2]);
        v3 bary = V3(0, 0, 0);
        int front = 1;
        int hit = intersectRayTriangle(q.o, q.d, v0, v1, v2, &bary);
        if (!hit) { front = 0; hit = intersectRayTriangle(q.o, q.d, v0, v2, v1, &bary); }
        if (hit && bary.z > 0.0f && bary.z < tmin) { tmin = bary.z; best = i; bestFront = front; }
    }
    if (best < 0) return -1;
    v3 v0 = V3(tris[best].v0[0], tris[best].v0[1], tris[best].v0[2]);
    v3 v1 = V3(tris[best].v1[0], tris[best].v1[1], tris[best].v1[2]);
    v3 v2 = V3(tris[best].v2[0], tris[best].v2[1], tris[best].v2[2]);
    v3 P = mulMV(mesh->transform, getPointOnRay(q, tmin), 1.0f);
    v3 n = v3normalize(mulMV(mesh->invTranspose, v3cross(v3sub(v1, v0), v3sub(v2, v0)), 0.0f));
    *normal = bestFront ? n : v3neg(n);
    *outside = bestFront;
    return v3length(v3sub(r.o, P));
}

/* src/pathtrace.cu:162-211: strictly smaller t wins => lowest index wins ties */
static int closestHit(const OrcGeom *geoms, int ngeoms, const OrcMeshes *meshes, ORay ray, float *tOut, v3 *nOut, int *outsideOut) {
    float t_min = FLT_MAX;
    int hit = -1, outside = 1, hitOutside = 1;
    v3 normal = V3(0, 0, 0), tmpN = V3(0, 0, 0);
    for (int i = 0; i < ngeoms; i++) {
        float t = -1.0f;
        if (geoms[i].type == 1) t = boxIntersectionTest(&geoms[i], ray, &tmpN, &outside);
        else if (geoms[i].type == 0) t = sphereIntersectionTest(&geoms[i], ray, &tmpN, &outside);
        else if (geoms[i].type == ORC_MESH && meshes)
            t = meshIntersectionTest(&geoms[i], meshes->triangles + meshes->ranges[2 * i], meshes->ranges[2 * i + 1], ray, &tmpN, &outside);
        if (t > 0.0f && t_min > t) {
            t_min = t; hit = i; normal = tmpN; hitOutside = outside;
        }
    }
    if (hit < 0) { *tOut = -1.0f; *nOut = V3(0, 0, 0); *outsideOut = 1; }
    else { *tOut = t_min; *nOut = normal; *outsideOut = hitOutside; }
    return hit;
}

/* src/pathtrace.cu:131-138 */
static inline ORay cameraRay(const OrcCamera *cam, int x, int y) {
    ORay r;
    r.o = V3(cam->position[0], cam->position[1], cam->position[2]);
    /* right * pl.x * (...)  ==  (right * pl.x) * (...)   -- left to right */
    v3 a = v3scale(v3scale(V3(cam->right[0], cam->right[1], cam->right[2]), cam->pixelLength[0]),
                   ((float)x - (float)cam->resolution[0] * 0.5f));
    v3 b = v3scale(v3scale(V3(cam->up[0], cam->up[1], cam->up[2]), cam->pixelLength[1]),
                   ((float)y - (float)cam->resolution[1] * 0.5f));
    r.d = v3normalize(v3sub(v3sub(V3(cam->view[0], cam->view[1], cam->view[2]), a), b));
    return r;
}

/* Camera ray with the OrcLens extensions (orc_api.h; the reference's TODOs at src/pathtrace.cu:118-120,134).
 * One engine per (iter, pixel) keyed by ORC_RAYGEN_DEPTH_KEY; draws, in order: jitter x, jitter y
 * (antialias only), lens radius, lens angle (lensRadius > 0 only).
 *   antialias : the sample position (x, y) becomes (x + u - 0.5, y + u' - 0.5) in src/pathtrace.cu:135-136
 *   thin lens : PBRT 6.2.3 -- the point in focus is position + dir * focalDistance / dot(dir, view); the
 *               origin moves to a uniformly sampled point of the lens disc spanned by right / up */
static inline ORay cameraRayLens(const OrcCamera *cam, const OrcLens *lens, int x, int y, int iter) {
    if (!lens || (!lens->antialias && !(lens->lensRadius > 0.0f))) return cameraRay(cam, x, y);
    Rng rng = seededEngine(iter, x + y * cam->resolution[0], ORC_RAYGEN_DEPTH_KEY);
    float sx = (float)x, sy = (float)y;
    if (lens->antialias) {
        sx = sx + (u01(&rng) - 0.5f);
        sy = sy + (u01(&rng) - 0.5f);
    }
    const v3 right = V3(cam->right[0], cam->right[1], cam->right[2]);
    const v3 up = V3(cam->up[0], cam->up[1], cam->up[2]);
    const v3 view = V3(cam->view[0], cam->view[1], cam->view[2]);
    ORay r;
    r.o = V3(cam->position[0], cam->position[1], cam->position[2]);
    v3 a = v3scale(v3scale(right, cam->pixelLength[0]), (sx - (float)cam->resolution[0] * 0.5f));
    v3 b = v3scale(v3scale(up, cam->pixelLength[1]), (sy - (float)cam->resolution[1] * 0.5f));
    r.d = v3normalize(v3sub(v3sub(view, a), b));
    if (lens->lensRadius > 0.0f) {
        float rad = lens->lensRadius * sqrtf(u01(&rng));
        float phi = u01(&rng) * TWO_PI_F;
        float lx = rad * cosf(phi), ly = rad * sinf(phi);
        float ft = lens->focalDistance / v3dot(r.d, view);
        v3 focus = v3add(r.o, v3scale(r.d, ft));
        r.o = v3add(v3add(r.o, v3scale(right, lx)), v3scale(up, ly));
        r.d = v3normalize(v3sub(focus, r.o));
    }
    return r;
}

/* the tangent frame and combination of src/interactions.h:21-41 for a given cos(theta) = up */
static v3 lobeDirection(v3 normal, float up, float over, float around) {
    v3 notNormal;
    if (fabsf(normal.x) < SQRT_OF_ONE_THIRD_F) notNormal = V3(1, 0, 0);
    else if (fabsf(normal.y) < SQRT_OF_ONE_THIRD_F) notNormal = V3(0, 1, 0);
    else notNormal = V3(0, 0, 1);
    v3 p1 = v3normalize(v3cross(normal, notNormal));
    v3 p2 = v3normalize(v3cross(normal, p1));
    v3 a = v3scale(normal, up);
    v3 b = v3scale(p1, cosf(around) * over);
    v3 c = v3scale(p2, sinf(around) * over);
    return v3add(v3add(a, b), c);
}

/* Imperfect specular (INSTRUCTION.md:202-216; GPU Gems 3 ch. 20, eq. 7-9): a direction around `axis`
 * (unit) with cos(theta) = u1^(1/(n+1)), phi = 2 pi u2 -- two draws, in that order. */
static v3 phongLobe(v3 axis, float exponent, Rng *rng) {
    float up = powf(u01(rng), 1.0f / (exponent + 1.0f));
    float s = 1 - up * up;
    float over = s > 0.0f ? sqrtf(s) : 0.0f;
    float around = u01(rng) * TWO_PI_F;
    return lobeDirection(axis, up, over, around);
}

/* src/interactions.h:10-42 */
static v3 hemisphere(v3 normal, Rng *rng) {
    float up = sqrtf(u01(rng));
    float over = sqrtf(1 - up * up);
    float around = u01(rng) * TWO_PI_F;
    v3 notNormal;
    if (fabsf(normal.x) < SQRT_OF_ONE_THIRD_F) notNormal = V3(1, 0, 0);
    else if (fabsf(normal.y) < SQRT_OF_ONE_THIRD_F) notNormal = V3(0, 1, 0);
    else notNormal = V3(0, 0, 1);
    v3 p1 = v3normalize(v3cross(normal, notNormal));
    v3 p2 = v3normalize(v3cross(normal, p1));
    v3 a = v3scale(normal, up);
    v3 b = v3scale(p1, cosf(around) * over);
    v3 c = v3scale(p2, sinf(around) * over);
    return v3add(v3add(a, b), c);
}

/* glm/detail/func_geometric.inl:176-179: I - N * dot(N, I) * 2 */
static inline v3 reflect3(v3 I, v3 N) {
    return v3sub(I, v3scale(v3scale(N, v3dot(N, I)), 2.0f));
}

/* the scatterRay body (src/interactions.h:44-79 is an empty TODO); spec: SURVEY.md section 9 */
static void scatter(ORay *ray, v3 *color, float t, v3 n, int outside, const OrcMaterial *m, int mode, Rng *rng) {
    v3 P = getPointOnRay(*ray, t);
    if (mode == ORC_MODE_FULL && m->hasRefractive > 0.0f) {
        v3 I = v3normalize(ray->d);
        float ior = m->indexOfRefraction;
        float eta = outside ? 1.0f / ior : ior;
        float cosi = -v3dot(n, I);
        float r0 = (1.0f - ior) / (1.0f + ior);
        r0 = r0 * r0;
        float c1 = 1.0f - cosi;
        float c2 = c1 * c1;
        float c5 = c2 * c2 * c1;
        float R = r0 + (1.0f - r0) * c5;
        float u = u01(rng);
        float dotNI = v3dot(n, I);
        float k = 1.0f - eta * eta * (1.0f - dotNI * dotNI);
        int tir = !(k >= 0.0f);
        if (tir || u < R) {
            ray->d = reflect3(I, n);
            ray->o = P;
            *color = v3mul(*color, V3(m->specColor[0], m->specColor[1], m->specColor[2]));
        } else {
            /* glm::refract (func_geometric.inl:193-200): (eta*I - (eta*dot + sqrt(k))*N) * float(k >= 0) */
            float f = eta * dotNI + sqrtf(k);
            v3 T = v3scale(v3sub(v3scale(I, eta), v3scale(n, f)), 1.0f);
            ray->o = getPointOnRay(*ray, t + 0.0002f);
            ray->d = T;
            *color = v3mul(*color, V3(m->color[0], m->color[1], m->color[2]));
        }
    } else if (mode == ORC_MODE_FULL && m->hasReflective > 0.0f) {
        v3 R = reflect3(ray->d, n);
        if (m->specExponent > 0.0f) {
            /* SPECEX n > 0: Phong lobe around the mirror direction; a sample that dips below the
             * surface is mirrored back above it */
            R = phongLobe(v3normalize(R), m->specExponent, rng);
            if (v3dot(R, n) < 0.0f) R = reflect3(R, n);
        }
        ray->d = R;
        ray->o = P;
        *color = v3mul(*color, V3(m->specColor[0], m->specColor[1], m->specColor[2]));
    } else {
        ray->d = hemisphere(n, rng);
        ray->o = P;
        *color = v3mul(*color, V3(m->color[0], m->color[1], m->color[2]));
    }
}

/* one shade step; returns 1 while alive */
static int shadeStep(const OrcMaterial *materials, int mode, int hit, int materialId, float t, v3 n, int outside,
                     int iter, int index, int depth, ORay *ray, v3 *color, int *remaining) {
    if (hit < 0) {                                      /* src/pathtrace.cu:262-264 */
        *color = V3(0, 0, 0);
        *remaining = 0;
        return 0;
    }
    const OrcMaterial *m = &materials[materialId];
    if (m->emittance > 0.0f) {                          /* src/pathtrace.cu:247-249 */
        v3 e = v3scale(V3(m->color[0], m->color[1], m->color[2]), m->emittance);
        *color = v3mul(*color, e);
        *remaining = 0;
        return 0;
    }
    Rng rng = seededEngine(iter, index, depth);         /* src/pathtrace.cu:240, keyed by pixelIndex */
    scatter(ray, color, t, n, outside, m, mode, &rng);
    (*remaining)--;
    if (*remaining == 0) {
        *color = V3(0, 0, 0);
        return 0;
    }
    return 1;
}

static inline v3 ld3(const float *p) { return V3(p[0], p[1], p[2]); }
static inline void st3(float *p, v3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

void orc_generate_rays(const OrcCamera *cam, float *origins, float *dirs) {
    const int W = cam->resolution[0], H = cam->resolution[1];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            ORay r = cameraRay(cam, x, y);
            size_t i = (size_t)x + (size_t)y * (size_t)W;
            st3(origins + 3 * i, r.o);
            st3(dirs + 3 * i, r.d);
        }
}

void orc_generate_rays_lens(const OrcCamera *cam, const OrcLens *lens, int iter, float *origins, float *dirs) {
    const int W = cam->resolution[0], H = cam->resolution[1];
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            ORay r = cameraRayLens(cam, lens, x, y, iter);
            size_t i = (size_t)x + (size_t)y * (size_t)W;
            st3(origins + 3 * i, r.o);
            st3(dirs + 3 * i, r.d);
        }
}

void orc_closest_hit(const OrcGeom *geoms, int ngeoms, const float *origins, const float *dirs,
                     int n, int *hit, float *t, float *normal, int *outside) {
    orc_closest_hit_mesh(geoms, ngeoms, NULL, origins, dirs, n, hit, t, normal, outside);
}

void orc_closest_hit_mesh(const OrcGeom *geoms, int ngeoms, const OrcMeshes *meshes,
                          const float *origins, const float *dirs, int n,
                          int *hit, float *t, float *normal, int *outside) {
#pragma omp parallel for schedule(dynamic, 1024)
    for (int i = 0; i < n; i++) {
        ORay r; r.o = ld3(origins + 3 * (size_t)i); r.d = ld3(dirs + 3 * (size_t)i);
        float tt; v3 nn; int oo;
        hit[i] = closestHit(geoms, ngeoms, meshes, r, &tt, &nn, &oo);
        t[i] = tt; st3(normal + 3 * (size_t)i, nn); outside[i] = oo;
    }
}

void orc_hemisphere(const float *normals, int n, const int *iter, const int *index,
                    const int *depth, float *outDirs) {
    for (int i = 0; i < n; i++) {
        Rng rng = seededEngine(iter[i], index[i], depth[i]);
        st3(outDirs + 3 * (size_t)i, hemisphere(ld3(normals + 3 * (size_t)i), &rng));
    }
}

void orc_shade(const OrcMaterial *materials, int mode, int n,
               const int *hit, const int *materialIds, const float *t,
               const float *normal, const int *outside,
               const int *iter, const int *index, const int *depth,
               float *origins, float *dirs, float *colors, int *remaining) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; i++) {
        ORay r; r.o = ld3(origins + 3 * (size_t)i); r.d = ld3(dirs + 3 * (size_t)i);
        v3 c = ld3(colors + 3 * (size_t)i);
        int rem = remaining[i];
        shadeStep(materials, mode, hit[i], hit[i] < 0 ? 0 : materialIds[i], t[i], ld3(normal + 3 * (size_t)i),
                  outside[i] != 0, iter[i], index[i], depth[i], &r, &c, &rem);
        st3(origins + 3 * (size_t)i, r.o); st3(dirs + 3 * (size_t)i, r.d); st3(colors + 3 * (size_t)i, c);
        remaining[i] = rem;
    }
}

long long orc_render(const OrcGeom *geoms, int ngeoms, const OrcMaterial *materials, int nmaterials,
                     const OrcCamera *cam, int traceDepth, int mode,
                     int iterBegin, int nIters, float *image, long long *alivePerDepth,
                     int nthreads) {
    return orc_render_ex(geoms, ngeoms, NULL, materials, nmaterials, cam, NULL, traceDepth, mode, iterBegin, nIters,
                         image, alivePerDepth, nthreads);
}

long long orc_render_ex(const OrcGeom *geoms, int ngeoms, const OrcMeshes *meshes,
                        const OrcMaterial *materials, int nmaterials,
                        const OrcCamera *cam, const OrcLens *lens, int traceDepth, int mode,
                        int iterBegin, int nIters, float *image, long long *alivePerDepth,
                        int nthreads) {
    (void)nmaterials;
    const int W = cam->resolution[0], H = cam->resolution[1];
    const int nd = traceDepth > 0 ? traceDepth : 1;
    long long segments = 0;
    long long *alive = (long long *)calloc((size_t)nd, sizeof(long long));
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        long long *myAlive = (long long *)calloc((size_t)nd, sizeof(long long));
        long long mySeg = 0;
#pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < H; y++) {
            for (int x = 0; x < W; x++) {
                const int index = x + y * W;                          /* src/pathtrace.cu:128 */
                for (int it = 0; it < nIters; it++) {
                    const int iter = iterBegin + it;
                    ORay ray = cameraRayLens(cam, lens, x, y, iter);
                    v3 color = V3(1.0f, 1.0f, 1.0f);                  /* src/pathtrace.cu:132 */
                    int remaining = traceDepth;                       /* src/pathtrace.cu:141 */
                    int depth = 0;
                    while (remaining > 0) {
                        myAlive[depth]++;
                        mySeg++;
                        float t; v3 n; int outside;
                        int hit = closestHit(geoms, ngeoms, meshes, ray, &t, &n, &outside);
                        int mid = hit < 0 ? 0 : geoms[hit].materialid;
                        int go = shadeStep(materials, mode, hit, mid, t, n, outside, iter, index, depth,
                                           &ray, &color, &remaining);
                        depth++;
                        if (!go) break;
                    }
                    image[3 * (size_t)index + 0] += color.x;          /* src/pathtrace.cu:276 */
                    image[3 * (size_t)index + 1] += color.y;
                    image[3 * (size_t)index + 2] += color.z;
                }
            }
        }
#pragma omp critical
        {
            for (int d = 0; d < nd; d++) alive[d] += myAlive[d];
            segments += mySeg;
        }
        free(myAlive);
    }
    if (alivePerDepth)
        for (int d = 0; d < traceDepth; d++) alivePerDepth[d] += alive[d];
    free(alive);
    return segments;
}

/* ------------------------------------------------------------------------------------------------
 * sin / cos restated.  The loop above calls libm's sinf / cosf, as the host build of the reference does
 * (src/interactions.h:38-39 resolve to the float overloads).  The CUDA path cannot call libm, so it
 * carries a restatement of glibc's algorithm (sysdeps/ieee754/flt-32/s_sincosf.h: double-precision
 * reduction by pi/2 and two double polynomials, one rounding to float).  The same restatement is kept
 * here so that it can be checked against THIS machine's libm over every float the samplers can produce
 * (orc_sincos_exhaustive) and used as the bit-exact expectation of the device function.
 * ------------------------------------------------------------------------------------------------ */
static void sincosRestated(float y, float *sn, float *cs) {
    const double x = (double)y;
    const int n = ((int)(x * 0x1.45F306DC9C883p+23) + 0x800000) >> 24;     /* reduce_fast, s_sincosf.h */
    const double r = x - (double)n * 0x1.921FB54442D18p0;
    const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2;
    const double c2 = -0x1.6c087e89a359dp-10 + r2 * 0x1.99343027bf8c3p-16;
    const double s1 = 0x1.1107605230bc4p-7 + r2 * -0x1.994eb3774cf24p-13;
    const double c1 = 1.0 + r2 * -0x1.ffffffd0c621cp-2;
    const double r5 = r3 * r2, r6 = r4 * r2;
    const double s = r + r3 * -0x1.555545995a603p-3;
    const double c = c1 + r4 * 0x1.55553e1068f19p-5;
    float S = (float)(s + r5 * s1), C = (float)(c + r6 * c2);
    if (n & 1) { const float t = S; S = C; C = -t; }
    if (n & 2) { S = -S; C = -C; }
    uint32_t u;
    memcpy(&u, &y, 4);
    const int tiny = ((u >> 20) & 0x7ffu) < 0x398u;                          /* |y| < 2^-12 */
    *sn = tiny ? y : S;
    *cs = tiny ? 1.0f : C;
}
void orc_sincos(const float *y, int n, int restated, float *sn, float *cs) {
    for (int i = 0; i < n; ++i) {
        if (restated) sincosRestated(y[i], &sn[i], &cs[i]);
        else { sn[i] = sinf(y[i]); cs[i] = cosf(y[i]); }
    }
}
/* number of floats with bit pattern in [loBits, hiBits] for which the restatement and libm disagree */
long long orc_sincos_exhaustive(unsigned loBits, unsigned hiBits, int nthreads) {
    long long bad = 0;
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long long b = (long long)loBits; b <= (long long)hiBits; ++b) {
        const uint32_t u = (uint32_t)b;
        float y, s, c;
        memcpy(&y, &u, 4);
        sincosRestated(y, &s, &c);
        const float ls = sinf(y), lc = cosf(y);
        bad += memcmp(&s, &ls, 4) != 0 || memcmp(&c, &lc, 4) != 0;
    }
    return bad;
}
