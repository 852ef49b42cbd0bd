// TEST INFRASTRUCTURE ONLY -- minimal stand-in for the OpenCV 3.4 types the reference's
// hot-path sources name (cv::cuda::GpuMat, cv::Ptr, cv::Rect, cv::Size, cv::Mat), so
// that core/src/TPS_RGBD*.cu and core/src/dense_registration*.cu compile UNMODIFIED,
// from where they lie under /root/reference, into the reference-kernel harness
// (oracle/_ref/libssf_ref.so).  OpenCV itself is not in this image.  Nothing here is
// reference code; semantics follow the public OpenCV API.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <memory>

#define CV_8U 0
#define CV_16U 2
#define CV_32S 4
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};

struct Rect {
  int x, y, width, height;
  __host__ __device__ Rect() : x(0), y(0), width(0), height(0) {}
  __host__ __device__ Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
  Size size() const { return Size(width, height); }
};

struct Mat {};

template <typename T>
struct Ptr {
  std::shared_ptr<T> p;
  Ptr() {}
  Ptr(T* q) : p(q) {}
  T* operator->() const { return p.get(); }
  T& operator*() const { return *p; }
  T* get() const { return p.get(); }
  bool empty() const { return !p; }
};

enum ColorConversionCodes { COLOR_BGR2BGRA = 0 };

namespace cuda {

inline size_t elem_size_of(int type) {
  const int depth = type & 7, cn = (type >> 3) + 1;
  const size_t d = (depth == CV_8U) ? 1 : (depth == CV_16U) ? 2 : 4;
  return d * cn;
}

struct GpuMat {
  unsigned char* data;
  size_t step;
  int rows, cols;
  int type_;
  std::shared_ptr<void> owner;
  GpuMat() : data(nullptr), step(0), rows(0), cols(0), type_(0) {}
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;
    void* ptr = nullptr;
    size_t pitch = 0;
    if (cudaMallocPitch(&ptr, &pitch, (size_t)c * elem_size_of(type), (size_t)r) != cudaSuccess) {
      fprintf(stderr, "ref shim: cudaMallocPitch failed\n");
      exit(-1);
    }
    cudaMemset2D(ptr, pitch, 0, (size_t)c * elem_size_of(type), (size_t)r);
    owner = std::shared_ptr<void>(ptr, [](void* q) { cudaFree(q); });
    data = static_cast<unsigned char*>(ptr);
    step = pitch; rows = r; cols = c; type_ = type;
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  Size size() const { return Size(cols, rows); }
  int type() const { return type_; }
  bool empty() const { return data == nullptr; }
  size_t elemSize() const { return elem_size_of(type_); }
  void download(Mat&) const {}
  // dense host <-> pitched device copies used by the harness driver
  void upload_dense(const void* host) {
    cudaMemcpy2D(data, step, host, (size_t)cols * elemSize(), (size_t)cols * elemSize(), rows, cudaMemcpyHostToDevice);
  }
  void download_dense(void* host) const {
    cudaMemcpy2D(host, (size_t)cols * elemSize(), data, step, (size_t)cols * elemSize(), rows, cudaMemcpyDeviceToHost);
  }
};

// 3 -> 4 channel copy, alpha = 255 (the only conversion the path uses, TPS_RGBD.cu:136)
void cvtColor(const GpuMat& src, GpuMat& dst, int code, int dcn = 0);

}  // namespace cuda
}  // namespace cv
