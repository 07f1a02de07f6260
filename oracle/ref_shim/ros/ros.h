/* shim: event_detector.cc includes <ros/ros.h> and uses nothing from it; the reference's
 * parameters.h declares two functions that take a ros::NodeHandle& (parameters.h:67-68). */
#pragma once
#include <string>
#include <vector>
namespace ros {
class NodeHandle;
}
