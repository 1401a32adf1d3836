// Stand-in for <ceres/loss_function.h>: see ../ssfm_mini_ceres.hpp (test infrastructure, NOT Ceres).
#pragma once
#include "../ssfm_mini_ceres.hpp"
