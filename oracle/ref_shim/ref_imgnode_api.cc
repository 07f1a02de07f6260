// ref_imgnode_api.cc -- TEST INFRASTRUCTURE: C entry points around the reference's own image node,
// feature_tracker/src/stereo_image_tracker_node.cpp, compiled UNMODIFIED (main() renamed on the
// command line) on top of the unmodified feature_tracker.cpp + event_detector.cc into
// oracle/_ref/libesvio_ref_imgnode.so.  What runs as the reference wrote it: img_callback_left/right
// (depth-1 queues, :36-52), sync_process (pairing, 1 s tolerance with its asymmetric comparisons,
// :216-258), getImageFromMsg (:261-317), handle_stereo_image (:55-183: first-frame skip, restart,
// publish-rate gate, the node's own CLAHE, trackImage, PointCloud packing).  Used only by
// tests/test_oracle_ref_node.py to pin esvio_b200/node.py's StereoImageNode / ImagePairer.
#include "feature_tracker.h"

#include <atomic>
#include <cstring>
#include <mutex>
#include <thread>

#include <cv_bridge/cv_bridge.h>

std::string IMAGE_TOPIC, IMAGE_LEFT, IMAGE_RIGHT, EVENT_TOPIC, EVENT_LEFT, EVENT_RIGHT, IMU_TOPIC, FISHEYE_MASK;
std::vector<std::string> CAM_NAMES;
void readParameters(ros::NodeHandle&) {}
void readParameters_event(ros::NodeHandle&) {}

extern ros::Publisher pub_img, pub_match, pub_restart;   // defined by the node file (:24-26)
extern FeatureTracker trackerData;
extern double first_image_time, last_image_time;
extern int pub_count;
extern bool first_image_flag, init_pub;
extern std::queue<sensor_msgs::ImageConstPtr> img_left_buf, img_right_buf;
extern std::mutex m_buf;
void handle_stereo_image(cv::Mat& img_left, cv::Mat& img_right, double msg_timestamp);
void img_callback_left(const sensor_msgs::ImageConstPtr&);
void img_callback_right(const sensor_msgs::ImageConstPtr&);
void sync_process();

void esvio_ref_apply_config(const int* cfg, const double* dcfg);          // ref_ft_api.cc
void esvio_ref_install_cameras(FeatureTracker& ft, const double* dcfg);

struct Cloud {
  uint32_t sec, nsec;
  std::vector<float> rows;
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

REF_API void ref_imgnode_reset(const int* cfg, const double* dcfg, int freq) {
  esvio_ref_apply_config(cfg, dcfg);
  FREQ = freq;
  trackerData.~FeatureTracker();
  new (&trackerData) FeatureTracker();
  esvio_ref_install_cameras(trackerData, dcfg);
  first_image_flag = true;
  first_image_time = last_image_time = 0;
  pub_count = 1;
  init_pub = false;
  PUB_THIS_FRAME = false;
  {
    std::lock_guard<std::mutex> lock(m_buf);
    while (!img_left_buf.empty()) img_left_buf.pop();
    while (!img_right_buf.empty()) img_right_buf.pop();
  }
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  g_clouds.clear();
  g_restarts = 0;
}

static cv::Mat mat_of(const uint8_t* img) {
  cv::Mat m = cv::Mat::zeros(cv::Size(COL, ROW), CV_8UC1);
  memcpy(m.ptr(), img, (size_t)COL * ROW);
  return m;
}

// handle_stereo_image(img_left, img_right, msg_timestamp) directly (:55); returns PUB_THIS_FRAME
REF_API int ref_imgnode_handle(const uint8_t* left, const uint8_t* right, double msg_timestamp) {
  cv::Mat l = mat_of(left), r = mat_of(right);
  handle_stereo_image(l, r, msg_timestamp);
  return PUB_THIS_FRAME ? 1 : 0;
}

// img_callback_left / _right: a mono8 sensor_msgs/Image arrives on a topic
REF_API void ref_imgnode_push_image(int cam, const uint8_t* img, double stamp) {
  std::shared_ptr<sensor_msgs::Image> m = std::make_shared<sensor_msgs::Image>();
  m->header.stamp = ros::Time(stamp);
  m->height = ROW, m->width = COL, m->step = COL;
  m->encoding = "mono8";
  m->data.assign(img, img + (size_t)COL * ROW);
  if (cam == 0) img_callback_left(m);
  else img_callback_right(m);
}

REF_API void ref_imgnode_start_sync_thread() {
  static std::atomic<bool> started{false};
  if (started.exchange(true)) return;
  std::thread(sync_process).detach();
}
// see ref_node_park_sync_thread (ref_node_api.cc)
REF_API void ref_imgnode_park_sync_thread() { m_buf.lock(); }
REF_API void ref_imgnode_queue_sizes(int* left, int* right) {
  std::lock_guard<std::mutex> lock(m_buf);
  *left = (int)img_left_buf.size();
  *right = (int)img_right_buf.size();
}
REF_API int ref_imgnode_n_clouds() {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  return (int)g_clouds.size();
}
REF_API int ref_imgnode_n_restarts() {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  return g_restarts;
}
REF_API int ref_imgnode_cloud_points(int i) {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  return (int)(g_clouds[i].rows.size() / 8);
}
REF_API void ref_imgnode_cloud(int i, uint32_t* sec, uint32_t* nsec, float* rows) {
  std::lock_guard<std::mutex> lock(g_cap_mutex);
  *sec = g_clouds[i].sec;
  *nsec = g_clouds[i].nsec;
  if (!g_clouds[i].rows.empty()) memcpy(rows, g_clouds[i].rows.data(), sizeof(float) * g_clouds[i].rows.size());
}
REF_API void ref_imgnode_state(double* first_time, double* last_time, int* count, int* first_flag, int* init) {
  *first_time = first_image_time;
  *last_time = last_image_time;
  *count = pub_count;
  *first_flag = first_image_flag ? 1 : 0;
  *init = init_pub ? 1 : 0;
}
REF_API double ref_imgnode_tracker_time() { return trackerData.cur_time; }
REF_API double ref_imgnode_tracker_prev_time() { return trackerData.prev_time; }
