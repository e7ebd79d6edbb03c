// norm.cuh — BatchNorm (train mode) kernels for the policy networks; filled in below.
#pragma once
#include "ptx.cuh"
namespace rovr {}
