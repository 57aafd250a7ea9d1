// TEST INFRASTRUCTURE (oracle/) -- /root/reference/common/include/fpga-power.h:5 includes
// nlohmann/json.hpp only for the XRT power sampler, which the oracle does not build.
#pragma once
namespace nlohmann { struct json {}; }
