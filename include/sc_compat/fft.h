/* fft.h -- drop-in for the reference header of the same name; see singlecarrier_compat.h */
#pragma once
#include "singlecarrier_compat.h"
