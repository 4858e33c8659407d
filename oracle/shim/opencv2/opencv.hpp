// TEST INFRASTRUCTURE ONLY (oracle tier A). Stand-in for the un-vendored OpenCV dependency of the
// reference (material.h:6, material.cpp:6, pathTracing.cpp:24): cv::Mat{rows, cols, empty(), at<Vec3b>},
// cv::Vec3b and cv::imread.  imread does not decode JPEG: it reads the side-car "<path>.bgr" written by
// tools/predecode_textures.py with Python cv2.imread (OpenCV's own decoder, BGR byte order):
//   bytes 0-3 "BGR8", int32 rows, int32 cols, rows*cols*3 bytes.   "parity unpinned" (no OpenCV pin).
#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

namespace cv
{
struct Vec3b
{
    unsigned char v[3];
    unsigned char &operator[](int i) { return v[i]; }
    const unsigned char &operator[](int i) const { return v[i]; }
};

class Mat
{
public:
    int rows = 0, cols = 0;
    unsigned char *data = nullptr; // cv::Mat::data: first byte of the (continuous) pixel buffer
    bool empty() const { return !buf || buf->empty(); }
    template <typename T>
    T &at(int r, int c) { return reinterpret_cast<T *>(buf->data())[(size_t)r * cols + c]; }
    std::shared_ptr<std::vector<unsigned char>> buf; // shared like cv::Mat's ref-counted header copy
};

inline Mat imread(const std::string &path)
{
    Mat m;
    FILE *f = fopen((path + ".bgr").c_str(), "rb");
    if (!f)
        return m;
    char magic[4];
    int32_t rc[2];
    if (fread(magic, 1, 4, f) == 4 && magic[0] == 'B' && magic[1] == 'G' && magic[2] == 'R' && magic[3] == '8' &&
        fread(rc, 4, 2, f) == 2 && rc[0] > 0 && rc[1] > 0)
    {
        auto b = std::make_shared<std::vector<unsigned char>>((size_t)rc[0] * rc[1] * 3);
        if (fread(b->data(), 1, b->size(), f) == b->size())
        {
            m.buf = b;
            m.data = b->data();
            m.rows = rc[0];
            m.cols = rc[1];
        }
    }
    fclose(f);
    return m;
}
} // namespace cv
