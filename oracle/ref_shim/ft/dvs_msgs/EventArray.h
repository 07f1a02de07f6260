// Stand-in for feature_tracker/src/dvs_msgs/EventArray.h (EventArray.h:44-54, fields only).
#pragma once
#include <std_msgs/Header.h>
#include "Event.h"
namespace dvs_msgs {
struct EventArray {
  typedef std::shared_ptr<const EventArray> ConstPtr;
  std_msgs::Header header;
  uint32_t height = 0, width = 0;
  std::vector<Event> events;
};
}  // namespace dvs_msgs
