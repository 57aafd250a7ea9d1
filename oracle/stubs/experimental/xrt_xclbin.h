// TEST INFRASTRUCTURE (oracle/) -- see xrt_device.h
#pragma once
#include <xrt/xrt_device.h>
