// Stand-in for rosbag (TEST INFRASTRUCTURE): EventMessageEditor.cpp writes its re-packed
// EventArray messages with Bag::write -- here the call hands the message to the harness
// (ref_windower_api.cc); reading bags (View / MessageInstance, used by the tool's main() only)
// is present so that the file compiles, and yields nothing.
#pragma once
#include <cstdlib>
#include <ros/ros.h>
#include <dvs_msgs/EventArray.h>
void esvio_ref_shim_bag_write(void* bag, const char* topic, const ros::Time& stamp, const dvs_msgs::EventArray& msg);
namespace rosbag {
namespace bagmode { enum BagMode { Write = 1, Read = 2, Append = 4 }; }
struct Bag {
  void open(const char*, int) {}
  bool isOpen() const { return false; }
  void close() {}
  void write(const char* topic, const ros::Time& stamp, const dvs_msgs::EventArray& msg) {
    esvio_ref_shim_bag_write(this, topic, stamp, msg);
  }
};
struct TopicQuery {
  TopicQuery(const std::string&) {}
};
struct MessageInstance {
  template <class T>
  std::shared_ptr<const T> instantiate() const { return std::shared_ptr<const T>(); }
};
struct View {
  View(Bag&, const TopicQuery&) {}
  const MessageInstance* begin() const { return nullptr; }
  const MessageInstance* end() const { return nullptr; }
};
}  // namespace rosbag
