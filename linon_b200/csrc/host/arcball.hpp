// arcball.hpp — Shoemake arcball camera (host mirror of src/arcball.rs:18-129 on cgmath 0.18 semantics,
// Cargo.lock:486-489).  fp32, column-major 4x4 matrices m[col][row].
#pragma once
#include <array>
#include <cmath>

namespace linon {

struct Mat4 {
    float m[4][4];
    static Mat4 identity() {
        Mat4 r{};
        for (int i = 0; i < 4; ++i) r.m[i][i] = 1.0f;
        return r;
    }
    static Mat4 from_translation(float x, float y, float z) {
        Mat4 r = identity();
        r.m[3][0] = x; r.m[3][1] = y; r.m[3][2] = z;
        return r;
    }
};

inline Mat4 operator*(const Mat4& a, const Mat4& b) {
    Mat4 r{};
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row) {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s = s + a.m[k][row] * b.m[c][k];
            r.m[c][row] = s;
        }
    return r;
}

inline std::array<float, 4> mul(const Mat4& a, const std::array<float, 4>& v) {
    std::array<float, 4> r{};
    for (int row = 0; row < 4; ++row) {
        float s = 0.0f;
        for (int k = 0; k < 4; ++k) s = s + a.m[k][row] * v[k];
        r[row] = s;
    }
    return r;
}

// adjugate / determinant (the structure of cgmath's Matrix4::invert)
inline bool invert(const Mat4& a, Mat4& out) {
    auto det3 = [&](int c0, int c1, int c2, int r0, int r1, int r2) {
        return a.m[c0][r0] * (a.m[c1][r1] * a.m[c2][r2] - a.m[c2][r1] * a.m[c1][r2]) -
               a.m[c1][r0] * (a.m[c0][r1] * a.m[c2][r2] - a.m[c2][r1] * a.m[c0][r2]) +
               a.m[c2][r0] * (a.m[c0][r1] * a.m[c1][r2] - a.m[c1][r1] * a.m[c0][r2]);
    };
    float cof[4][4];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            int cs[3], rs[3], ci = 0, ri = 0;
            for (int k = 0; k < 4; ++k) {
                if (k != c) cs[ci++] = k;
                if (k != r) rs[ri++] = k;
            }
            const float sign = ((c + r) & 1) ? -1.0f : 1.0f;
            cof[c][r] = sign * det3(cs[0], cs[1], cs[2], rs[0], rs[1], rs[2]);
        }
    float det = 0.0f;
    for (int c = 0; c < 4; ++c) det = det + a.m[c][0] * cof[c][0];
    if (det == 0.0f) return false;
    const float inv_det = 1.0f / det;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out.m[c][r] = cof[r][c] * inv_det;
    return true;
}

struct Quat {
    float s, x, y, z;
};
inline Quat operator*(const Quat& a, const Quat& b) {  // Hamilton product
    return {a.s * b.s - a.x * b.x - a.y * b.y - a.z * b.z, a.s * b.x + a.x * b.s + a.y * b.z - a.z * b.y,
            a.s * b.y + a.y * b.s + a.z * b.x - a.x * b.z, a.s * b.z + a.z * b.s + a.x * b.y - a.y * b.x};
}
inline Mat4 to_mat4(const Quat& q) {  // cgmath: impl From<Quaternion> for Matrix4
    const float x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
    const float xx2 = x2 * q.x, xy2 = x2 * q.y, xz2 = x2 * q.z;
    const float yy2 = y2 * q.y, yz2 = y2 * q.z, zz2 = z2 * q.z;
    const float sy2 = y2 * q.s, sz2 = z2 * q.s, sx2 = x2 * q.s;
    Mat4 r{};
    r.m[0][0] = 1.0f - yy2 - zz2; r.m[0][1] = xy2 + sz2; r.m[0][2] = xz2 - sy2;
    r.m[1][0] = xy2 - sz2; r.m[1][1] = 1.0f - xx2 - zz2; r.m[1][2] = yz2 + sx2;
    r.m[2][0] = xz2 + sy2; r.m[2][1] = yz2 - sx2; r.m[2][2] = 1.0f - xx2 - yy2;
    r.m[3][3] = 1.0f;
    return r;
}

enum class CameraOperation { None, Rotate, Pan };  // src/arcball.rs:11-16

class ArcballCamera {
public:
    std::array<float, 3> center;

    ArcballCamera(std::array<float, 3> center_, float zoom_speed, float screen_w, float screen_h)
        : center(center_), zoom_speed_(zoom_speed) {  // src/arcball.rs:33-46
        translation_ = Mat4::from_translation(0.0f, 0.0f, -1.0f);
        invert(Mat4::from_translation(center[0], center[1], center[2]), center_translation_);
        rotation_ = {1.0f, 0.0f, 0.0f, 0.0f};
        inv_screen_[0] = 1.0f / screen_w;
        inv_screen_[1] = 1.0f / screen_h;
        update_camera();
    }
    const Mat4& get_mat4() const { return camera_; }
    std::array<float, 3> eye_pos() const { return {inv_camera_.m[3][0], inv_camera_.m[3][1], inv_camera_.m[3][2]}; }
    std::array<float, 3> eye_dir() const { return direction({0.0f, 0.0f, -1.0f, 0.0f}); }
    std::array<float, 3> up_dir() const { return direction({0.0f, 1.0f, 0.0f, 0.0f}); }

    void rotate(float prev_x, float prev_y, float cur_x, float cur_y) {  // src/arcball.rs:73-88
        auto clamp1 = [](float v) { return v < -1.0f ? -1.0f : (v > 1.0f ? 1.0f : v); };
        const float cx = clamp1(cur_x * 2.0f * inv_screen_[0] - 1.0f), cy = clamp1(1.0f - 2.0f * cur_y * inv_screen_[1]);
        const float px = clamp1(prev_x * 2.0f * inv_screen_[0] - 1.0f), py = clamp1(1.0f - 2.0f * prev_y * inv_screen_[1]);
        rotation_ = screen_to_arcball(cx, cy) * screen_to_arcball(px, py) * rotation_;
        update_camera();
    }
    void zoom(float amount, float elapsed) {  // src/arcball.rs:90-95
        const float dz = amount * zoom_speed_ * elapsed;
        translation_ = Mat4::from_translation(0.0f * zoom_speed_ * elapsed, 0.0f * zoom_speed_ * elapsed, dz) * translation_;
        update_camera();
    }
    void pan(float dx, float dy) {  // src/arcball.rs:97-110
        const float zoom_dist = std::fabs(translation_.m[3][3]);
        const std::array<float, 4> delta = {dx * inv_screen_[0] * zoom_dist, -dy * inv_screen_[1] * zoom_dist,
                                            0.0f * zoom_dist, 0.0f * zoom_dist};
        const std::array<float, 4> motion = mul(inv_camera_, delta);
        center_translation_ = Mat4::from_translation(motion[0], motion[1], motion[2]) * center_translation_;
        update_camera();
    }
    void update_screen(float w, float h) {  // src/arcball.rs:112-115
        inv_screen_[0] = 1.0f / w;
        inv_screen_[1] = 1.0f / h;
    }

private:
    Mat4 translation_, center_translation_, camera_, inv_camera_;
    Quat rotation_;
    float zoom_speed_;
    float inv_screen_[2];

    void update_camera() {  // src/arcball.rs:116-119
        camera_ = translation_ * to_mat4(rotation_) * center_translation_;
        invert(camera_, inv_camera_);
    }
    std::array<float, 3> direction(const std::array<float, 4>& v) const {
        const std::array<float, 4> d = mul(inv_camera_, v);
        const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        return {d[0] / len, d[1] / len, d[2] / len};
    }
    static Quat screen_to_arcball(float x, float y) {  // src/arcball.rs:120-129
        const float dist = x * x + y * y;
        if (dist <= 1.0f) return {0.0f, x, y, std::sqrt(1.0f - dist)};
        const float len = std::sqrt(dist);
        return {0.0f, x / len, y / len, 0.0f};
    }
};

}  // namespace linon
