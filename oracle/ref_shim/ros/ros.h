/* shim: event_detector.cc / feature_tracker.cpp include <ros/ros.h>; they use the logging macros
 * only, and the reference's parameters.h declares two functions that take a ros::NodeHandle&
 * (parameters.h:67-68). */
#pragma once
#include <string>
#include <vector>
namespace ros {
class NodeHandle;
}
#define ROS_DEBUG(...) ((void)0)
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
