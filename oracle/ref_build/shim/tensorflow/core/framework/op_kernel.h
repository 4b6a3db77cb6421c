// Stand-in for the handful of TensorFlow framework symbols that the reference's CPU plan
// (tensorflow_nufft/cc/kernels/nufft_plan.{h,cc}) touches. TEST INFRASTRUCTURE ONLY: lets
// oracle/ref_build compile the reference sources where they lie, unmodified. Not TF code.
#pragma once
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <initializer_list>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
namespace tensorflow {
using std::string;
class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(std::string m) : ok_(false), msg_(std::move(m)) {}
  bool ok() const { return ok_; }
  const std::string& message() const { return msg_; }
 private:
  bool ok_;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
namespace errors {
template <typename... A> Status Make(const char* kind, A&&... a) {
  std::ostringstream s; s << kind << ": "; (s << ... << a); return Status(s.str());
}
template <typename... A> Status InvalidArgument(A&&... a) { return Make("InvalidArgument", a...); }
template <typename... A> Status Unimplemented(A&&... a) { return Make("Unimplemented", a...); }
template <typename... A> Status Internal(A&&... a) { return Make("Internal", a...); }
template <typename... A> Status ResourceExhausted(A&&... a) { return Make("ResourceExhausted", a...); }
}  // namespace errors
#define TF_RETURN_IF_ERROR(x) do { ::tensorflow::Status _s = (x); if (!_s.ok()) return _s; } while (0)
struct FatalLog {
  std::ostringstream s;
  template <typename T> FatalLog& operator<<(const T& t) { s << t; return *this; }
  ~FatalLog() { std::cerr << s.str() << std::endl; std::abort(); }
};
#define FATAL 0
#define LOG(x) ::tensorflow::FatalLog()
enum DataType { DT_FLOAT, DT_DOUBLE, DT_INT32, DT_COMPLEX64, DT_COMPLEX128 };
template <typename T> struct DataTypeToEnum;
template <> struct DataTypeToEnum<float> { static constexpr DataType value = DT_FLOAT; };
template <> struct DataTypeToEnum<double> { static constexpr DataType value = DT_DOUBLE; };
template <> struct DataTypeToEnum<int> { static constexpr DataType value = DT_INT32; };
template <> struct DataTypeToEnum<std::complex<float>> { static constexpr DataType value = DT_COMPLEX64; };
template <> struct DataTypeToEnum<std::complex<double>> { static constexpr DataType value = DT_COMPLEX128; };
inline size_t DataTypeSize(DataType d) {
  switch (d) { case DT_FLOAT: return 4; case DT_DOUBLE: return 8; case DT_INT32: return 4;
               case DT_COMPLEX64: return 8; default: return 16; }
}
struct TensorShape {
  int64_t n = 0;
  TensorShape() {}
  TensorShape(std::initializer_list<int64_t> l) { n = 1; for (auto v : l) n *= v; }
  int64_t num_elements() const { return n; }
};
template <typename T> struct Flat { T* p; int64_t n; T* data() { return p; } int64_t size() const { return n; } };
class Tensor {
 public:
  void Alloc(DataType d, int64_t n) {
    n_ = n;
    size_t bytes = ((static_cast<size_t>(n) * DataTypeSize(d) + 63) / 64) * 64 + 64;
    buf_ = std::shared_ptr<void>(std::aligned_alloc(64, bytes), std::free);
  }
  template <typename T> Flat<T> flat() { return Flat<T>{reinterpret_cast<T*>(buf_.get()), n_}; }
 private:
  std::shared_ptr<void> buf_;
  int64_t n_ = 0;
};
// ---- GPU-side pieces named by tensorflow_nufft_b200/tf_glue/nufft_kernels_b200.cc (syntax check of
// the glue only; signatures follow TF 2.11: framework/allocator.h, framework/device_base.h) ----
struct AllocatorAttributes {};
class Allocator {
 public:
  static constexpr size_t kAllocatorAlignment = 64;
  virtual ~Allocator() {}
  virtual void* AllocateRaw(size_t alignment, size_t num_bytes) { return std::aligned_alloc(alignment, (num_bytes + alignment - 1) / alignment * alignment); }
  virtual void DeallocateRaw(void* ptr) { std::free(ptr); }
};
class DeviceBase {
 public:
  struct CpuWorkerThreads { int num_threads = 1; };
  struct AcceleratorDeviceInfo { int gpu_id = 0; };
  const CpuWorkerThreads* tensorflow_cpu_worker_threads() const { return &cpu_; }
  const AcceleratorDeviceInfo* tensorflow_accelerator_device_info() const { return &acc_; }
  Allocator* GetAllocator(AllocatorAttributes) { return &alloc_; }
 private:
  CpuWorkerThreads cpu_;
  AcceleratorDeviceInfo acc_;
  Allocator alloc_;
};
class OpKernelContext {
 public:
  Status allocate_temp(DataType d, const TensorShape& s, Tensor* t) { t->Alloc(d, s.num_elements()); return OkStatus(); }
  template <typename D> const D& eigen_device() const { static D d; return d; }
  DeviceBase* device() const { return const_cast<DeviceBase*>(&dev_); }
 private:
  DeviceBase dev_;
};
}  // namespace tensorflow
