// ref_node_api.cc -- TEST INFRASTRUCTURE: C entry points around the reference's own event node,
// feature_tracker/src/stereo_event_tracker_node.cpp, compiled UNMODIFIED (with -Dmain=... so that
// its main() does not collide) together with feature_tracker.cpp and event_detector.cc into
// oracle/_ref/libesvio_ref_node.so (recipe: oracle/Makefile).  What runs as the reference wrote
// it: event_callback_left/right (the depth-1 queues, :128-142), sync_process (pairing with the
// 0.2 s tolerance, :372-419, on a thread of its own as in the node), handle_stereo_event
// (:145-344: empty-window and first-window skips, restart on a time jump, the publish-rate gate,
// the motion-compensation assembly from the IMU / odometry queues, trackEvent, the PointCloud
// packing, the suppressed first publish), imu_callback / state_callback.  ROS itself is a
// stand-in (oracle/ref_shim/ros/ros.h, ft/*_msgs): Publisher::publish hands the message to this
// file.  Used only by tests/test_oracle_ref_node.py to pin esvio_b200/node.py (and through the
// twin-trace test include/esvio_fe_node.hpp) on the reference's node code.
#include "feature_tracker.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

// ---- what parameters.cpp / visualization.cpp define and the node file references
std::string IMAGE_TOPIC, IMAGE_LEFT, IMAGE_RIGHT, EVENT_TOPIC, EVENT_LEFT, EVENT_RIGHT, IMU_TOPIC, FISHEYE_MASK;
std::vector<std::string> CAM_NAMES;
ros::Publisher pub_loop_image, pub_img, pub_match, pub_match_two, pub_time_surface, pub_restart, corner_pub,
    pub_match_two_point, pub_event_loop;
void registerPub(ros::NodeHandle&) {}
void readParameters(ros::NodeHandle&) {}
void readParameters_event(ros::NodeHandle&) {}

// ---- the node file's globals and functions (stereo_event_tracker_node.cpp:27-62,104-142,145,372)
extern FeatureTracker trackerData;
extern double first_image_time, last_image_time, last_imu_t, t_pre, t_cur;
extern int pub_count;
extern bool first_image_flag, init_pub, is_nolinear;
extern Eigen::Vector3f v_cur, v_pre;
extern std::queue<sensor_msgs::ImuConstPtr> imu_buf;
extern std::queue<nav_msgs::Odometry::ConstPtr> odom_buffer_;
extern std::queue<dvs_msgs::EventArray> events_left_buf, events_right_buf;
extern std::mutex m_buf_event;
void handle_stereo_event(const dvs_msgs::EventArray&, const dvs_msgs::EventArray&, double);
void event_callback_left(const dvs_msgs::EventArray&);
void event_callback_right(const dvs_msgs::EventArray&);
void imu_callback(const sensor_msgs::ImuConstPtr&);
void state_callback(const nav_msgs::Odometry::ConstPtr&);
void sync_process();

void esvio_ref_apply_config(const int* cfg, const double* dcfg);          // ref_ft_api.cc
void esvio_ref_install_cameras(FeatureTracker& ft, const double* dcfg);

// ---- captured publications
struct Cloud {
  uint32_t sec, nsec;
  std::vector<float> rows;  // 8 per point: x y z | id*2+cam, u, v, vx, vy
};
static std::mutex g_cap_mutex;
static std::vector<Cloud> g_clouds;
static int g_restarts = 0;

void esvio_ref_shim_publish(const ros::Publisher* pub, const std::type_info& type, const void* msg) {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  if (pub == &pub_restart && type == typeid(std_msgs::Bool)) {
    if (static_cast<const std_msgs::Bool*>(msg)->data) ++g_restarts;
  } else if (pub == &pub_img && type == typeid(sensor_msgs::PointCloudPtr)) {
    const sensor_msgs::PointCloud& pc = **static_cast<const sensor_msgs::PointCloudPtr*>(msg);
    Cloud c;
    c.sec = pc.header.stamp.sec;
    c.nsec = pc.header.stamp.nsec;
    const size_t n = pc.points.size();
    c.rows.resize(8 * n);
    for (size_t i = 0; i < n; ++i) {
      float* r = &c.rows[8 * i];
      r[0] = pc.points[i].x, r[1] = pc.points[i].y, r[2] = pc.points[i].z;
      for (int ch = 0; ch < 5; ++ch) r[3 + ch] = pc.channels[ch].values[i];
    }
    g_clouds.push_back(c);
  }
}

#define REF_API extern "C" __attribute__((visibility("default")))

// same cfg / dcfg as ref_ft_create (ref_ft_api.cc); the node's own state as a fresh process has it
REF_API void ref_node_reset(const int* cfg, const double* dcfg, int freq, int do_motion_correction) {
  esvio_ref_apply_config(cfg, dcfg);
  FREQ = freq;
  Do_motion_correction = do_motion_correction;
  trackerData.~FeatureTracker();
  new (&trackerData) FeatureTracker();
  esvio_ref_install_cameras(trackerData, dcfg);
  first_image_flag = true;
  first_image_time = last_image_time = last_imu_t = 0;
  pub_count = 1;
  init_pub = false;
  is_nolinear = false;
  v_cur = Eigen::Vector3f();
  v_pre = Eigen::Vector3f();
  t_pre = t_cur = 0;
  PUB_THIS_FRAME = false;
  while (!imu_buf.empty()) imu_buf.pop();
  while (!odom_buffer_.empty()) odom_buffer_.pop();
  {
    std::lock_guard<std::mutex> lock(m_buf_event);
    while (!events_left_buf.empty()) events_left_buf.pop();
    while (!events_right_buf.empty()) events_right_buf.pop();
  }
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  g_clouds.clear();
  g_restarts = 0;
}

static void fill(dvs_msgs::EventArray& a, const uint16_t* x, const uint16_t* y, const uint32_t* sec,
                 const uint32_t* nsec, const uint8_t* pol, size_t n, uint32_t st_sec, uint32_t st_nsec) {
  a.events.resize(n);
  for (size_t i = 0; i < n; ++i) {
    dvs_msgs::Event& e = a.events[i];
    e.x = x[i], e.y = y[i], e.ts.sec = sec[i], e.ts.nsec = nsec[i], e.polarity = pol[i];
  }
  a.header.stamp.sec = st_sec;
  a.header.stamp.nsec = st_nsec;
}

// handle_stereo_event(event_left, event_right, msg_timestamp) directly (node.cpp:145); returns
// PUB_THIS_FRAME as the call left it
REF_API int ref_node_handle(const uint16_t* lx, const uint16_t* ly, const uint32_t* lsec, const uint32_t* lnsec,
                            const uint8_t* lp, size_t nl, const uint16_t* rx, const uint16_t* ry,
                            const uint32_t* rsec, const uint32_t* rnsec, const uint8_t* rp, size_t nr,
                            uint32_t stamp_sec, uint32_t stamp_nsec, double msg_timestamp) {
  dvs_msgs::EventArray L, R;
  fill(L, lx, ly, lsec, lnsec, lp, nl, stamp_sec, stamp_nsec);
  fill(R, rx, ry, rsec, rnsec, rp, nr, stamp_sec, stamp_nsec);
  std::streambuf* old = std::cout.rdbuf(nullptr);  // detector.init(..., fx, ...) prints the matrix
  handle_stereo_event(L, R, msg_timestamp);
  std::cout.rdbuf(old);
  return PUB_THIS_FRAME ? 1 : 0;
}

// event_callback_left / _right (node.cpp:128-142): one message arrives on a topic
REF_API void ref_node_push_events(int cam, const uint16_t* x, const uint16_t* y, const uint32_t* sec,
                                  const uint32_t* nsec, const uint8_t* pol, size_t n, uint32_t stamp_sec,
                                  uint32_t stamp_nsec) {
  dvs_msgs::EventArray a;
  fill(a, x, y, sec, nsec, pol, n, stamp_sec, stamp_nsec);
  if (cam == 0) event_callback_left(a);
  else event_callback_right(a);
}

REF_API void ref_node_push_imu(double t, const double* omega, const double* accel) {
  std::shared_ptr<sensor_msgs::Imu> m = std::make_shared<sensor_msgs::Imu>();
  m->header.stamp = ros::Time(t);
  m->angular_velocity.x = omega[0], m->angular_velocity.y = omega[1], m->angular_velocity.z = omega[2];
  m->linear_acceleration.x = accel[0], m->linear_acceleration.y = accel[1], m->linear_acceleration.z = accel[2];
  imu_callback(m);
}

REF_API void ref_node_push_odometry(double t, const double* v) {
  std::shared_ptr<nav_msgs::Odometry> m = std::make_shared<nav_msgs::Odometry>();
  m->header.stamp = ros::Time(t);
  m->twist.twist.linear.x = v[0], m->twist.twist.linear.y = v[1], m->twist.twist.linear.z = v[2];
  state_callback(m);
}

// sync_process (node.cpp:372-419) on a thread of its own, as std::thread sync_thread{sync_process}
// in main() (:366); it never returns, so the thread is started once per process and detached
REF_API void ref_node_start_sync_thread() {
  static std::atomic<bool> started{false};
  if (started.exchange(true)) return;
  std::thread(sync_process).detach();
}

// sync_process never returns; before the process ends the harness parks it: the node's own mutex is
// taken and kept, so the thread blocks at its next poll and touches nothing while the library's
// statics are torn down.  Nothing that takes m_buf_event may be called afterwards.
REF_API void ref_node_park_sync_thread() { m_buf_event.lock(); }

// sizes of the two depth-1 queues (under the node's own mutex)
REF_API void ref_node_queue_sizes(int* left, int* right) {
  std::lock_guard<std::mutex> lock(m_buf_event);
  *left = (int)events_left_buf.size();
  *right = (int)events_right_buf.size();
}

REF_API int ref_node_n_clouds() {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  return (int)g_clouds.size();
}
REF_API int ref_node_n_restarts() {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  return g_restarts;
}
REF_API int ref_node_cloud_points(int i) {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  return (int)(g_clouds[i].rows.size() / 8);
}
REF_API void ref_node_cloud(int i, uint32_t* sec, uint32_t* nsec, float* rows) {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  *sec = g_clouds[i].sec;
  *nsec = g_clouds[i].nsec;
  if (!g_clouds[i].rows.empty()) memcpy(rows, g_clouds[i].rows.data(), sizeof(float) * g_clouds[i].rows.size());
}
// node state the mirror keeps too (for the decision trace)
REF_API void ref_node_state(double* first_time, double* last_time, int* count, int* first_flag, int* init) {
  *first_time = first_image_time;
  *last_time = last_image_time;
  *count = pub_count;
  *first_flag = first_image_flag ? 1 : 0;
  *init = init_pub ? 1 : 0;
}
// the clock handed to trackEvent last (FeatureTracker::cur_time): node.cpp:190,193,254
REF_API double ref_node_tracker_time() { return trackerData.cur_time; }
// == cur_time once trackEvent has returned (feature_tracker.cpp:587)
REF_API double ref_node_tracker_prev_time() { return trackerData.prev_time; }
