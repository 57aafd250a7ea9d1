// TEST INFRASTRUCTURE (oracle/) -- empty stand-ins for the XRT types named by
// /root/reference/common/src/spmv-helper.cpp:936-1053 (fpgaRun) so the host packer and
// cpuSequential compile without Xilinx XRT.  fpgaRun itself is never called by the oracle.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#define XCL_BO_SYNC_BO_TO_DEVICE 0
#define XCL_BO_SYNC_BO_FROM_DEVICE 1
namespace xrt {
namespace info { enum class device { name, bdf, electrical }; }
struct uuid {};
struct device {
  device() = default;
  explicit device(int) {}
  explicit device(const std::string&) {}
  template <info::device>
  std::string get_info() const { return std::string(); }
  uuid load_xclbin(const std::string&) { return uuid(); }
};
struct kernel {
  enum class cu_access_mode { exclusive, shared };
  kernel() = default;
  kernel(const device&, const uuid&, const std::string&, cu_access_mode = cu_access_mode::shared) {}
  int group_id(int) const { return 0; }
};
struct run {
  run() = default;
  explicit run(const kernel&) {}
  template <typename T>
  void set_arg(int, T&&) {}
  void start() {}
  void wait() {}
};
struct bo {
  bo() = default;
  template <typename... A>
  bo(A&&...) {}
  void sync(int) {}
  void sync(int, size_t, size_t) {}
  template <typename T>
  T map() { return nullptr; }
};
}  // namespace xrt
