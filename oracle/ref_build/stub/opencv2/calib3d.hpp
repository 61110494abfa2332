// TEST INFRASTRUCTURE ONLY -- see opencv.hpp next to this file.
#pragma once
#include "opencv.hpp"
