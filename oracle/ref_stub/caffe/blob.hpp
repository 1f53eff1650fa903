// oracle/ref_stub -- TEST INFRASTRUCTURE.  Stand-in for caffe-mnc/include/caffe/blob.hpp: an
// (N,C,H,W) array with `data` and `diff`, host and device views.  Managed memory keeps the two
// views identical without SyncedMemory's lazy copies; on a box without a GPU the blob falls back
// to plain host memory (only the reference's Forward_cpu paths can run there).
#ifndef MNC_REF_STUB_BLOB_HPP_
#define MNC_REF_STUB_BLOB_HPP_
#include "caffe/common.hpp"

namespace caffe {
template <typename Dtype>
class Blob {
 public:
  Blob() : data_(0), diff_(0), count_(0), capacity_(0), managed_(false) {}
  ~Blob() { release(); }
  void Reshape(const int num, const int channels, const int height, const int width) {
    vector<int> s(4);
    s[0] = num; s[1] = channels; s[2] = height; s[3] = width;
    Reshape(s);
  }
  void Reshape(const vector<int>& shape) {
    shape_ = shape;
    count_ = 1;
    for (size_t i = 0; i < shape.size(); ++i) count_ *= shape[i];
    if (count_ > capacity_) {
      release();
      capacity_ = count_;
    }
  }
  const vector<int>& shape() const { return shape_; }
  int shape(int i) const { return shape_[i]; }
  int count() const { return count_; }
  int num() const { return dim(0); }
  int channels() const { return dim(1); }
  int height() const { return dim(2); }
  int width() const { return dim(3); }
  int offset(const int n, const int c = 0, const int h = 0, const int w = 0) const {
    return ((n * channels() + c) * height() + h) * width() + w;
  }
  const Dtype* cpu_data() const { sync(); return buf(&data_); }
  const Dtype* gpu_data() const { return buf(&data_); }
  const Dtype* cpu_diff() const { sync(); return buf(&diff_); }
  const Dtype* gpu_diff() const { return buf(&diff_); }
  Dtype* mutable_cpu_data() { sync(); return buf(&data_); }
  Dtype* mutable_gpu_data() { return buf(&data_); }
  Dtype* mutable_cpu_diff() { sync(); return buf(&diff_); }
  Dtype* mutable_gpu_diff() { return buf(&diff_); }

 private:
  Blob(const Blob&);
  Blob& operator=(const Blob&);
  int dim(size_t i) const { return i < shape_.size() ? shape_[i] : 1; }
  void sync() const { if (managed_) cudaDeviceSynchronize(); }
  Dtype* buf(Dtype* const* slot) const {
    Dtype** s = const_cast<Dtype**>(slot);
    if (*s == 0 && capacity_ > 0) {
      void* p = 0;
      size_t bytes = sizeof(Dtype) * static_cast<size_t>(capacity_);
      if (cudaMallocManaged(&p, bytes) == cudaSuccess) {
        managed_ = true;
        cudaMemset(p, 0, bytes);
        cudaDeviceSynchronize();
      } else {
        cudaGetLastError();
        p = std::calloc(bytes, 1);
      }
      *s = static_cast<Dtype*>(p);
    }
    return *s;
  }
  void release() {
    Dtype** slots[2] = {&data_, &diff_};
    for (int i = 0; i < 2; ++i) {
      if (*slots[i]) {
        if (managed_) cudaFree(*slots[i]); else std::free(*slots[i]);
        *slots[i] = 0;
      }
    }
  }
  Dtype* data_;
  Dtype* diff_;
  vector<int> shape_;
  int count_;
  int capacity_;
  mutable bool managed_;
};
}  // namespace caffe
#endif
