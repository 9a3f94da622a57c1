// stand-in header: see pagk_cv_shim.hpp
#pragma once
#include "../../pagk_cv_shim.hpp"
