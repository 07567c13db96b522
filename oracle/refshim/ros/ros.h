/* reference-build shim: the ORB front end only uses ROS for assertions and logging */
#ifndef ORB_REFSHIM_ROS_H
#define ORB_REFSHIM_ROS_H
#include <cstdio>
#include <stdexcept>
#define ROS_ASSERT(cond) do { if (!(cond)) throw std::runtime_error("ROS_ASSERT failed: " #cond); } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)
#endif
