// oracle stub: the reference's parameters.h includes OpenCV for cv::FileStorage (parameters.cpp only, not compiled here)
#pragma once
