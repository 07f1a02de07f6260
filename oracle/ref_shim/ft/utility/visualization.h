// Stand-in for feature_tracker/src/utility/visualization.h (ROS publishers / markers).  The
// tracking path uses from it: the publisher objects the node publishes on
// (visualization.h:24-33) and registerPub; feature_tracker.h:238-244 names sensor_msgs::Imu.
#pragma once
#include <ros/ros.h>
#include <nav_msgs/Odometry.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/Imu.h>
#include <sensor_msgs/PointCloud.h>
#include <std_msgs/Bool.h>
#include <std_msgs/Header.h>
extern ros::Publisher pub_loop_image;
extern ros::Publisher pub_img, pub_match, pub_match_two;
extern ros::Publisher pub_time_surface;
extern ros::Publisher pub_restart;
extern ros::Publisher corner_pub;
extern ros::Publisher pub_match_two_point;
extern ros::Publisher pub_event_loop;
void registerPub(ros::NodeHandle& n);
