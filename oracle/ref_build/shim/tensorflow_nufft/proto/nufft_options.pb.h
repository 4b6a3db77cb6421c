// Hand-written stand-in for the protoc output of tensorflow_nufft/proto/nufft_options.proto
// (fields: debugging.check_points_range, fftw.planning_rigor, max_batch_size, points_range).
// Test infrastructure only.
#pragma once
namespace tensorflow { namespace nufft {
enum FftwPlanningRigor : int { AUTO = 0, ESTIMATE = 1, MEASURE = 2, PATIENT = 3, EXHAUSTIVE = 4 };
enum PointsRange : int { STRICT = 0, EXTENDED = 1, INFINITE = 2 };
class FftwOptions {
 public:
  FftwPlanningRigor planning_rigor() const { return r_; }
  void set_planning_rigor(FftwPlanningRigor r) { r_ = r; }
 private:
  FftwPlanningRigor r_ = AUTO;
};
class DebuggingOptions {
 public:
  bool check_points_range() const { return c_; }
  void set_check_points_range(bool c) { c_ = c; }
 private:
  bool c_ = false;
};
class Options {
 public:
  const DebuggingOptions& debugging() const { return d_; }
  DebuggingOptions* mutable_debugging() { return &d_; }
  const FftwOptions& fftw() const { return f_; }
  FftwOptions* mutable_fftw() { return &f_; }
  int max_batch_size() const { return b_; }
  void set_max_batch_size(int b) { b_ = b; }
  PointsRange points_range() const { return p_; }
  void set_points_range(PointsRange p) { p_ = p; }
 private:
  DebuggingOptions d_;
  FftwOptions f_;
  int b_ = 0;
  PointsRange p_ = STRICT;
};
}}  // namespace tensorflow::nufft
