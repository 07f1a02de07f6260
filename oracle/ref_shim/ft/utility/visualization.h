// Stand-in for feature_tracker/src/utility/visualization.h (ROS publishers / markers / tf; none of
// it is on the tracking path).  What the compiled reference files need from it: the message types
// they name, the publisher objects the nodes publish on, and registerPub.  TEST INFRASTRUCTURE.
#pragma once
#include <ros/ros.h>
#include <nav_msgs/Odometry.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/Imu.h>
#include <sensor_msgs/PointCloud.h>
#include <std_msgs/Bool.h>
#include <std_msgs/Header.h>

// one object per topic the reference advertises; the harness (ref_node_api.cc) defines them and
// recognises pub_img / pub_restart by address
#define ESVIO_SHIM_TOPIC(object) extern ros::Publisher object
ESVIO_SHIM_TOPIC(pub_img);              // "feature": the PointCloud for esvio_estimator
ESVIO_SHIM_TOPIC(pub_restart);          // "restart"
ESVIO_SHIM_TOPIC(pub_match);            // visualisation topics from here on
ESVIO_SHIM_TOPIC(pub_match_two);
ESVIO_SHIM_TOPIC(pub_match_two_point);
ESVIO_SHIM_TOPIC(pub_time_surface);
ESVIO_SHIM_TOPIC(pub_event_loop);
ESVIO_SHIM_TOPIC(pub_loop_image);
ESVIO_SHIM_TOPIC(corner_pub);
#undef ESVIO_SHIM_TOPIC

void registerPub(ros::NodeHandle& node_handle);
