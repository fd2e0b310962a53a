#include "opencv2/cv_minimal.hpp"
