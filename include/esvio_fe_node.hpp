// esvio_fe_node.hpp -- header-only, ROS-free mirror of the reference's stereo_event_tracker node
// around the tracker call, to be used with esvio::GpuFeatureTracker
// (esvio_fe_adapter.hpp) or any tracker type with the reference's member names:
//
//   EventWindower    dependences/events_repacking_helper/src/EventMessageEditor.cpp:8-57
//   EventPairer      feature_tracker/src/stereo_event_tracker_node.cpp:128-142, 372-419
//   MotionAssembler  stereo_event_tracker_node.cpp:102-125, 195-252
//   StereoEventNode  stereo_event_tracker_node.cpp:145-344
//
// `EventArrayT` is any type with `.events` (contiguous 16-byte dvs_msgs::Event records with
// `.ts.sec/.ts.nsec`) and `.stamp` (double seconds: header.stamp.toSec()).
#ifndef ESVIO_FE_NODE_HPP
#define ESVIO_FE_NODE_HPP

#include <cmath>
#include <deque>
#include <set>
#include <utility>
#include <vector>

#include "esvio_fe.h"

namespace esvio {

struct RosTime {  // ros::Time(double) / toSec()
  uint32_t sec, nsec;
  static RosTime fromSec(double t) {
    int64_t s = (int64_t)std::floor(t);
    int64_t ns = (int64_t)std::llround((t - (double)s) * 1e9);
    s += ns / 1000000000;
    ns %= 1000000000;
    return RosTime{(uint32_t)s, (uint32_t)ns};
  }
  double toSec() const { return (double)sec + 1e-9 * (double)nsec; }
};

template <class EventT>
inline double event_time(const EventT& e) {
  return (double)e.ts.sec + 1e-9 * (double)e.ts.nsec;
}

// EventMessageEditor: fixed-rate re-windowing of a raw event stream.  `Sink(EventArrayT&&)`
// receives every completed message (header stamp = the window's end time).
template <class EventArrayT>
class EventWindower {
 public:
  explicit EventWindower(double frequency = 30.0) : duration_(1.0 / frequency) {}
  template <class EventT, class Sink>
  void insertEvent(const EventT& e, Sink&& sink) {
    if (first_) {
      reset(event_time(e));
      first_ = false;
    }
    if (event_time(e) >= end_) {
      cur_.stamp = end_;
      sink(std::move(cur_));
      reset(end_);
    }
    cur_.events.push_back(e);
  }

 private:
  void reset(double start) {
    end_ = RosTime::fromSec(start + duration_).toSec();
    cur_ = EventArrayT();
    cur_.stamp = end_;
  }
  double duration_, end_ = 0.0;
  bool first_ = true;
  EventArrayT cur_;
};

// depth-1 queues (a new message replaces the waiting one) + the pairing step of sync_process
template <class EventArrayT>
class EventPairer {
 public:
  void pushLeft(EventArrayT m) { push(left_, std::move(m)); }
  void pushRight(EventArrayT m) { push(right_, std::move(m)); }
  bool ready() const { return !left_.empty() && !right_.empty(); }
  // one iteration of sync_process; true when a pair came out
  bool poll(EventArrayT* l, EventArrayT* r, double* msg_timestamp) {
    if (!ready()) return false;
    const double tl = left_.front().stamp, tr = right_.front().stamp;
    if (tl < tr - 0.2) {
      left_.pop_front();
      return false;
    }
    if (tl > tr + 0.2) {
      right_.pop_front();
      return false;
    }
    *msg_timestamp = tl;
    *l = std::move(left_.front());
    left_.pop_front();
    *r = std::move(right_.front());
    right_.pop_front();
    return true;
  }
  int dropped = 0;

 private:
  void push(std::deque<EventArrayT>& q, EventArrayT m) {
    if (!q.empty()) {
      q.pop_front();
      ++dropped;
    }
    q.push_back(std::move(m));
  }
  std::deque<EventArrayT> left_, right_;
};

struct ImuSample {
  double stamp;
  double wx, wy, wz;
};
struct OdomSample {
  double stamp;
  double vx, vy, vz;
};

// the Motion_correction_value of one window from the IMU / odometry queues (node.cpp:195-252)
class MotionAssembler {
 public:
  bool pushImu(const ImuSample& m) {
    if (m.stamp <= last_imu_t_) return false;  // "imu message in disorder!"
    last_imu_t_ = m.stamp;
    imu_.push_back(m);
    return true;
  }
  void pushOdometry(const OdomSample& m) { odom_.push_back(m); }
  esvio_motion assemble(double t_left_0, double t_left_1) {
    esvio_motion mc{};
    if (!imu_.empty()) {
      if (!odom_.empty()) {
        const OdomSample o = odom_.front();
        odom_.pop_front();
        mc.state_v[0] = o.vx, mc.state_v[1] = o.vy, mc.state_v[2] = o.vz;
        for (int i = 0; i < 3; ++i) v_pre_[i] = v_cur_[i];
        v_cur_[0] = (float)o.vx, v_cur_[1] = (float)o.vy, v_cur_[2] = (float)o.vz;
        t_pre_ = t_cur_;
        t_cur_ = o.stamp;
        for (int i = 0; i < 3; ++i) mc.accel[i] = (float)((v_cur_[i] - v_pre_[i]) / (t_cur_ - t_pre_));
      }
      while (!imu_.empty() && imu_.front().stamp < t_left_0) imu_.pop_front();
      if (!imu_.empty()) {
        mc.omega[0] = (float)imu_.front().wx;
        mc.omega[1] = (float)imu_.front().wy;
        mc.omega[2] = (float)imu_.front().wz;
      }
    }
    for (int i = 0; i < 3; ++i) mc.v_pre[i] = v_pre_[i];
    mc.t1 = t_left_1;
    return mc;
  }

 private:
  std::deque<ImuSample> imu_;
  std::deque<OdomSample> odom_;
  double last_imu_t_ = 0.0, t_pre_ = 0.0, t_cur_ = 0.0;
  float v_cur_[3] = {0, 0, 0}, v_pre_[3] = {0, 0, 0};
};

struct CloudRow {  // points[i] = (x, y, 1); channels = {id*2+cam, u, v, vx, vy}
  float x, y, z, id_cam, u, v, vx, vy;
};
struct FeatureCloud {
  double stamp = 0.0;  // header.stamp = ros::Time(msg_timestamp), frame_id "world"
  std::vector<CloudRow> rows;
};

// handle_stereo_event.  TrackerT: trackEvent(double, const EventArrayT&, const EventArrayT&),
// optionally trackEvent(double, L, R, const esvio_motion&), PUB_THIS_FRAME and the public
// result vectors of feature_tracker.h:126-135.
template <class TrackerT>
class StereoEventNode {
 public:
  StereoEventNode(TrackerT& tracker, int freq, bool do_motion_correction = false)
      : t_(tracker), freq_(freq), do_mc_(do_motion_correction) {}
  MotionAssembler motion;
  int restarts = 0, windows_tracked = 0;

  // true when `cloud` was published for this pair
  template <class EventArrayT>
  bool handle_stereo_event(const EventArrayT& event_left, const EventArrayT& event_right,
                           double msg_timestamp, FeatureCloud* cloud) {
    if (event_left.events.empty()) return false;
    if (first_image_flag_) {
      first_image_flag_ = false;
      first_image_time_ = msg_timestamp;
      last_image_time_ = msg_timestamp;
      return false;
    }
    if (msg_timestamp - last_image_time_ > 1.0 || msg_timestamp < last_image_time_) {
      first_image_flag_ = true;
      last_image_time_ = 0;
      pub_count_ = 1;
      ++restarts;
      return false;
    }
    last_image_time_ = msg_timestamp;
    bool pub;
    if (std::round(1.0 * pub_count_ / (msg_timestamp - first_image_time_)) <= freq_) {
      pub = true;
      if (std::fabs(1.0 * pub_count_ / (msg_timestamp - first_image_time_) - freq_) < 0.01 * freq_) {
        first_image_time_ = msg_timestamp;
        pub_count_ = 0;
      }
    } else {
      pub = false;
    }
    t_.PUB_THIS_FRAME = pub;
    const double t_last = event_time(event_left.events.back());
    if (!do_mc_) {
      t_.trackEvent(t_last, event_left, event_right);
    } else {
      const esvio_motion mc = motion.assemble(event_time(event_left.events.front()), event_left.stamp);
      track_mc(t_last, msg_timestamp, event_left, event_right, mc, 0);
    }
    ++windows_tracked;
    if (!pub) return false;
    ++pub_count_;
    cloud->stamp = msg_timestamp;
    cloud->rows.clear();
    std::set<int> hash_ids;
    for (size_t j = 0; j < t_.ids.size(); ++j)
      if (t_.track_cnt[j] > 1) {
        hash_ids.insert(t_.ids[j]);
        cloud->rows.push_back({t_.cur_un_pts[j].x, t_.cur_un_pts[j].y, 1.f,
                               (float)(t_.ids[j] * 2 + 0), t_.cur_pts[j].x, t_.cur_pts[j].y,
                               t_.pts_velocity[j].x, t_.pts_velocity[j].y});
      }
    for (size_t j = 0; j < t_.ids_right.size(); ++j)
      if (hash_ids.count(t_.ids_right[j]))
        cloud->rows.push_back({t_.cur_un_right_pts[j].x, t_.cur_un_right_pts[j].y, 1.f,
                               (float)(t_.ids_right[j] * 2 + 1), t_.cur_right_pts[j].x,
                               t_.cur_right_pts[j].y, t_.right_pts_velocity[j].x,
                               t_.right_pts_velocity[j].y});
    if (!init_pub_) {  // the first cloud is never published (node.cpp:334-339)
      init_pub_ = true;
      return false;
    }
    return true;
  }

 private:
  template <class EA, class T = TrackerT>
  auto track_mc(double t_last, double, const EA& l, const EA& r, const esvio_motion& mc, int)
      -> decltype(std::declval<T&>().trackEvent(0.0, l, r, mc), void()) {
    // stereo_event_tracker_node.cpp:190,254: msg_timestamp_left = events.back().ts, the same
    // clock as the plain call; the header stamp only travels inside mc.t1 (:197,202)
    t_.trackEvent(t_last, l, r, mc);
  }
  template <class EA>
  void track_mc(double t_last, double, const EA& l, const EA& r, const esvio_motion&, long) {
    t_.trackEvent(t_last, l, r);
  }
  TrackerT& t_;
  int freq_;
  bool do_mc_;
  bool first_image_flag_ = true, init_pub_ = false;
  double first_image_time_ = 0.0, last_image_time_ = 0.0;
  int pub_count_ = 1;
};

// The depth-1 queues of img_callback_left/right (stereo_image_tracker_node.cpp:36-52) and the
// pairing step of the image node's sync_process (:217-241): +-1 s; a left frame exactly one
// second older than the right one is already thrown (`<=`, unlike the event node's `<`).
// MsgT needs a `double stamp`.
template <class MsgT>
class ImagePairer {
 public:
  void pushLeft(MsgT m) { push(left_, std::move(m)); }
  void pushRight(MsgT m) { push(right_, std::move(m)); }
  bool ready() const { return !left_.empty() && !right_.empty(); }
  bool poll(MsgT* l, MsgT* r, double* msg_timestamp) {
    if (!ready()) return false;
    const double tl = left_.front().stamp, tr = right_.front().stamp;
    if (tl <= tr - 1) {
      left_.pop_front();
      return false;
    }
    if (tl > tr + 1) {
      right_.pop_front();
      return false;
    }
    *msg_timestamp = tl;
    *l = std::move(left_.front());
    left_.pop_front();
    *r = std::move(right_.front());
    right_.pop_front();
    return true;
  }
  int dropped = 0;

 private:
  void push(std::deque<MsgT>& q, MsgT m) {
    if (!q.empty()) {
      q.pop_front();
      ++dropped;
    }
    q.push_back(std::move(m));
  }
  std::deque<MsgT> left_, right_;
};

// handle_stereo_image (stereo_image_tracker_node.cpp:54-183): the same first-frame skip,
// restart rule, publish-rate gate, cloud packing and first-publish suppression as the event
// node around trackerData.trackImage(msg_timestamp, img_left, img_right) (:99).  The node's
// own CLAHE (EQUALIZE, :93-97) is not mirrored.
template <class TrackerT>
class StereoImageNode {
 public:
  StereoImageNode(TrackerT& tracker, int freq) : t_(tracker), freq_(freq) {}
  int restarts = 0, frames_tracked = 0;

  template <class MatT>
  bool handle_stereo_image(const MatT& img_left, const MatT& img_right, double msg_timestamp,
                           FeatureCloud* cloud) {
    if (first_image_flag_) {
      first_image_flag_ = false;
      first_image_time_ = msg_timestamp;
      last_image_time_ = msg_timestamp;
      return false;
    }
    if (msg_timestamp - last_image_time_ > 1.0 || msg_timestamp < last_image_time_) {
      first_image_flag_ = true;
      last_image_time_ = 0;
      pub_count_ = 1;
      ++restarts;
      return false;
    }
    last_image_time_ = msg_timestamp;
    bool pub;
    if (std::round(1.0 * pub_count_ / (msg_timestamp - first_image_time_)) <= freq_) {
      pub = true;
      if (std::fabs(1.0 * pub_count_ / (msg_timestamp - first_image_time_) - freq_) < 0.01 * freq_) {
        first_image_time_ = msg_timestamp;
        pub_count_ = 0;
      }
    } else {
      pub = false;
    }
    t_.PUB_THIS_FRAME = pub;
    t_.trackImage(msg_timestamp, img_left, img_right);
    ++frames_tracked;
    if (!pub) return false;
    ++pub_count_;
    cloud->stamp = msg_timestamp;
    cloud->rows.clear();
    std::set<int> hash_ids;
    for (size_t j = 0; j < t_.ids.size(); ++j)
      if (t_.track_cnt[j] > 1) {
        hash_ids.insert(t_.ids[j]);
        cloud->rows.push_back({t_.cur_un_pts[j].x, t_.cur_un_pts[j].y, 1.f,
                               (float)(t_.ids[j] * 2 + 0), t_.cur_pts[j].x, t_.cur_pts[j].y,
                               t_.pts_velocity[j].x, t_.pts_velocity[j].y});
      }
    for (size_t j = 0; j < t_.ids_right.size(); ++j)
      if (hash_ids.count(t_.ids_right[j]))
        cloud->rows.push_back({t_.cur_un_right_pts[j].x, t_.cur_un_right_pts[j].y, 1.f,
                               (float)(t_.ids_right[j] * 2 + 1), t_.cur_right_pts[j].x,
                               t_.cur_right_pts[j].y, t_.right_pts_velocity[j].x,
                               t_.right_pts_velocity[j].y});
    if (!init_pub_) {  // :172-177
      init_pub_ = true;
      return false;
    }
    return true;
  }

 private:
  TrackerT& t_;
  int freq_;
  bool first_image_flag_ = true, init_pub_ = false;
  double first_image_time_ = 0.0, last_image_time_ = 0.0;
  int pub_count_ = 1;
};

}  // namespace esvio
#endif
