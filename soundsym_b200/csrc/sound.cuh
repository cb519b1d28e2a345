// sound.cuh — entry points of the MFCC / segmentation / resynthesis kernels (sound.cu, segment.cu)
#pragma once
#include "common.cuh"
