// TEST INFRASTRUCTURE ONLY -- stand-in for the subset of assimp's value types
// that the reference hot path touches (assimp@a5a53433042ae9f9f7d2f9d25312b46fd09702a4,
// pinned by /root/reference/vendor/assimp/CMakeLists.txt:19-20; not vendored,
// not installed, no network). Used ONLY to compile oracle/_ref from the
// reference's own sources. The arithmetic below restates assimp's published
// vector3.inl / color4.inl / matrix3x3.inl / matrix4x4.inl operator order
// (fp32, left-to-right sums). Reference call sites:
//   pathtracer.cpp:55-56,68-77,84,88,100-101  main.cpp:53-59,131,212-216
//   lib/types.h:92-123.
#pragma once
#include <cmath>
#include <cstring>

typedef float ai_real;

struct aiVector3D {
    float x, y, z;
    aiVector3D() : x(0), y(0), z(0) {}
    aiVector3D(float _x, float _y, float _z) : x(_x), y(_y), z(_z) {}
    bool operator==(const aiVector3D& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator!=(const aiVector3D& o) const { return !(*this == o); }
    float operator[](unsigned i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](unsigned i) { return i == 0 ? x : (i == 1 ? y : z); }
    float Length() const { return std::sqrt(x * x + y * y + z * z); }
};
inline aiVector3D operator+(const aiVector3D& a, const aiVector3D& b) { return aiVector3D(a.x + b.x, a.y + b.y, a.z + b.z); }
inline aiVector3D operator-(const aiVector3D& a, const aiVector3D& b) { return aiVector3D(a.x - b.x, a.y - b.y, a.z - b.z); }
inline aiVector3D operator*(float f, const aiVector3D& v) { return aiVector3D(f * v.x, f * v.y, f * v.z); }
inline aiVector3D operator*(const aiVector3D& v, float f) { return aiVector3D(f * v.x, f * v.y, f * v.z); }
// dot product
inline float operator*(const aiVector3D& a, const aiVector3D& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// cross product
inline aiVector3D operator^(const aiVector3D& a, const aiVector3D& b) {
    return aiVector3D(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct aiColor3D {
    float r, g, b;
    aiColor3D() : r(0), g(0), b(0) {}
    aiColor3D(float _r, float _g, float _b) : r(_r), g(_g), b(_b) {}
};

struct aiColor4D {
    float r, g, b, a;
    aiColor4D() : r(0), g(0), b(0), a(0) {}
    aiColor4D(float _r, float _g, float _b, float _a) : r(_r), g(_g), b(_b), a(_a) {}
    const aiColor4D& operator+=(const aiColor4D& o) { r += o.r; g += o.g; b += o.b; a += o.a; return *this; }
    const aiColor4D& operator-=(const aiColor4D& o) { r -= o.r; g -= o.g; b -= o.b; a -= o.a; return *this; }
    const aiColor4D& operator*=(float f) { r *= f; g *= f; b *= f; a *= f; return *this; }
    const aiColor4D& operator/=(float f) { r /= f; g /= f; b /= f; a /= f; return *this; }
    bool operator==(const aiColor4D& o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
    bool operator!=(const aiColor4D& o) const { return !(*this == o); }
};
inline aiColor4D operator+(const aiColor4D& v1, const aiColor4D& v2) { return aiColor4D(v1.r + v2.r, v1.g + v2.g, v1.b + v2.b, v1.a + v2.a); }
inline aiColor4D operator-(const aiColor4D& v1, const aiColor4D& v2) { return aiColor4D(v1.r - v2.r, v1.g - v2.g, v1.b - v2.b, v1.a - v2.a); }
inline aiColor4D operator*(const aiColor4D& v1, const aiColor4D& v2) { return aiColor4D(v1.r * v2.r, v1.g * v2.g, v1.b * v2.b, v1.a * v2.a); }
inline aiColor4D operator*(float f, const aiColor4D& v) { return aiColor4D(f * v.r, f * v.g, f * v.b, f * v.a); }
inline aiColor4D operator*(const aiColor4D& v, float f) { return aiColor4D(f * v.r, f * v.g, f * v.b, f * v.a); }
inline aiColor4D operator/(const aiColor4D& v, float f) { return v * (1 / f); }

struct aiMatrix4x4 {
    float a1, a2, a3, a4, b1, b2, b3, b4, c1, c2, c3, c4, d1, d2, d3, d4;
    aiMatrix4x4()
        : a1(1), a2(0), a3(0), a4(0), b1(0), b2(1), b3(0), b4(0), c1(0), c2(0), c3(1), c4(0), d1(0), d2(0), d3(0), d4(1) {}
};
inline aiVector3D operator*(const aiMatrix4x4& m, const aiVector3D& v) {
    aiVector3D res;
    res.x = m.a1 * v.x + m.a2 * v.y + m.a3 * v.z + m.a4;
    res.y = m.b1 * v.x + m.b2 * v.y + m.b3 * v.z + m.b4;
    res.z = m.c1 * v.x + m.c2 * v.y + m.c3 * v.z + m.c4;
    return res;
}

struct aiMatrix3x3 {
    float a1, a2, a3, b1, b2, b3, c1, c2, c3;
    aiMatrix3x3() : a1(1), a2(0), a3(0), b1(0), b2(1), b3(0), c1(0), c2(0), c3(1) {}
    explicit aiMatrix3x3(const aiMatrix4x4& m)
        : a1(m.a1), a2(m.a2), a3(m.a3), b1(m.b1), b2(m.b2), b3(m.b3), c1(m.c1), c2(m.c2), c3(m.c3) {}
    float* operator[](unsigned i) { return &a1 + 3 * i; }
    const float* operator[](unsigned i) const { return &a1 + 3 * i; }
    float Determinant() const {
        return a1 * b2 * c3 - a1 * b3 * c2 + a2 * b3 * c1 - a2 * b1 * c3 + a3 * b1 * c2 - a3 * b2 * c1;
    }
    aiMatrix3x3& Inverse() {
        float det = Determinant();
        if (det == 0.0f) {
            float nan = std::nanf("");
            a1 = a2 = a3 = b1 = b2 = b3 = c1 = c2 = c3 = nan;
            return *this;
        }
        float invdet = 1.0f / det;
        aiMatrix3x3 res;
        res.a1 = invdet * (b2 * c3 - b3 * c2);
        res.a2 = -invdet * (a2 * c3 - a3 * c2);
        res.a3 = invdet * (a2 * b3 - a3 * b2);
        res.b1 = -invdet * (b1 * c3 - b3 * c1);
        res.b2 = invdet * (a1 * c3 - a3 * c1);
        res.b3 = -invdet * (a1 * b3 - a3 * b1);
        res.c1 = invdet * (b1 * c2 - b2 * c1);
        res.c2 = -invdet * (a1 * c2 - a2 * c1);
        res.c3 = invdet * (a1 * b2 - a2 * b1);
        *this = res;
        return *this;
    }
    // Moeller & Hughes, "Efficiently building a matrix to rotate one vector to
    // another" (JGT 1999), as shipped in assimp's matrix3x3.inl.
    static aiMatrix3x3& FromToMatrix(const aiVector3D& from, const aiVector3D& to, aiMatrix3x3& mtx) {
        const float e = from * to;
        const float f = (e < 0) ? -e : e;
        if (f > 1.0f - 0.00001f) {
            aiVector3D u, v, x;
            x.x = (from.x > 0.0f) ? from.x : -from.x;
            x.y = (from.y > 0.0f) ? from.y : -from.y;
            x.z = (from.z > 0.0f) ? from.z : -from.z;
            if (x.x < x.y) {
                if (x.x < x.z) { x.x = 1.0f; x.y = x.z = 0.0f; }
                else { x.z = 1.0f; x.x = x.y = 0.0f; }
            } else {
                if (x.y < x.z) { x.y = 1.0f; x.x = x.z = 0.0f; }
                else { x.z = 1.0f; x.x = x.y = 0.0f; }
            }
            u.x = x.x - from.x; u.y = x.y - from.y; u.z = x.z - from.z;
            v.x = x.x - to.x; v.y = x.y - to.y; v.z = x.z - to.z;
            const float c1 = 2.0f / (u * u);
            const float c2 = 2.0f / (v * v);
            const float c3 = c1 * c2 * (u * v);
            for (unsigned i = 0; i < 3; i++) {
                for (unsigned j = 0; j < 3; j++) {
                    mtx[i][j] = -c1 * u[i] * u[j] - c2 * v[i] * v[j] + c3 * v[i] * u[j];
                }
                mtx[i][i] += 1.0f;
            }
        } else {
            const aiVector3D v = from ^ to;
            const float h = 1.0f / (1.0f + e);
            const float hvx = h * v.x;
            const float hvz = h * v.z;
            const float hvxy = hvx * v.y;
            const float hvxz = hvx * v.z;
            const float hvyz = hvz * v.y;
            mtx[0][0] = e + hvx * v.x;
            mtx[0][1] = hvxy - v.z;
            mtx[0][2] = hvxz + v.y;
            mtx[1][0] = hvxy + v.z;
            mtx[1][1] = e + h * v.y * v.y;
            mtx[1][2] = hvyz - v.x;
            mtx[2][0] = hvxz - v.y;
            mtx[2][1] = hvyz + v.x;
            mtx[2][2] = e + hvz * v.z;
        }
        return mtx;
    }
};
inline aiVector3D operator*(const aiMatrix3x3& m, const aiVector3D& v) {
    aiVector3D res;
    res.x = m.a1 * v.x + m.a2 * v.y + m.a3 * v.z;
    res.y = m.b1 * v.x + m.b2 * v.y + m.b3 * v.z;
    res.z = m.c1 * v.x + m.c2 * v.y + m.c3 * v.z;
    return res;
}

struct aiString {
    char data[64];
    aiString() { data[0] = 0; }
    const char* C_Str() const { return data; }
};

struct aiRay {
    aiVector3D pos, dir;
};
