/* Stand-in for roscpp (TEST INFRASTRUCTURE, see oracle/Makefile).  event_detector.cc and
 * feature_tracker.cpp use the logging macros only; stereo_event_tracker_node.cpp
 * (_ref/libesvio_ref_node.so) additionally names ros::Time, NodeHandle / Subscriber /
 * Publisher, init / spin.  What carries data here: ros::Time (fromSec / toSec as roscpp's
 * time.h implements them) and Publisher::publish, which hands the message to the harness. */
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <queue>
#include <set>
#include <string>
#include <typeinfo>
#include <vector>

namespace ros {
struct Time {
  uint32_t sec = 0, nsec = 0;
  Time() {}
  Time(uint32_t s, uint32_t ns) : sec(s), nsec(ns) {}
  /* TimeBase::fromSec: floor, nsec rounded to nearest, carry */
  explicit Time(double t) {
    const int64_t sec64 = (int64_t)std::floor(t);
    sec = (uint32_t)sec64;
    nsec = (uint32_t)std::llround((t - (double)sec) * 1e9);
    sec += (nsec / 1000000000ul);
    nsec %= 1000000000ul;
  }
  /* TimeBase::toSec */
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
  static Time now() { return Time(); }
};
struct TransportHints {
  TransportHints& tcpNoDelay() { return *this; }
};
struct Subscriber {};
struct Publisher;
}  // namespace ros
/* defined by the harness (ref_node_api.cc); other libraries never publish */
void esvio_ref_shim_publish(const ros::Publisher* pub, const std::type_info& type, const void* msg);
namespace ros {
struct Publisher {
  template <class M>
  void publish(const M& m) const { esvio_ref_shim_publish(this, typeid(M), &m); }
};
class NodeHandle {
 public:
  NodeHandle() {}
  NodeHandle(const char*) {}
  template <class F>
  Subscriber subscribe(const std::string&, int, F, TransportHints = TransportHints()) { return Subscriber(); }
  template <class M>
  Publisher advertise(const std::string&, int) { return Publisher(); }
};
inline void init(int&, char**, const char*) {}
inline void spin() {}
namespace console {
namespace levels { enum Level { Debug, Info, Warn, Error }; }
inline bool set_logger_level(const char*, levels::Level) { return true; }
}  // namespace console
}  // namespace ros
#define ROSCONSOLE_DEFAULT_NAME "ros"
#define ROS_DEBUG(...) ((void)0)
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_DEBUG_STREAM(x) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
