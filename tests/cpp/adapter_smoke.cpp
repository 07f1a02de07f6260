// Compiles the header-only adapter against a stand-in for dvs_msgs::EventArray (no ROS here) and
// drives a few windows through it.  Exit code 0: tracked features came back and the PointCloud
// rows keep the consumer's invariants; 3: no CUDA device (the expected outcome on a CPU box --
// there is no CPU fallback); anything else is a failure.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "esvio_fe_adapter.hpp"

namespace dvs_msgs {  // layout of feature_tracker/src/dvs_msgs/Event.h:42-52
struct Time {
  uint32_t sec, nsec;
};
struct Event {
  uint16_t x, y;
  Time ts;
  uint8_t polarity;
};
struct EventArray {
  std::vector<Event> events;
};
}  // namespace dvs_msgs
static_assert(sizeof(dvs_msgs::Event) == 16, "Event layout");

static uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// a square of side 40 px moving diagonally, seen by both cameras with 6 px disparity
static void make_window(int k, int cam, dvs_msgs::EventArray& out) {
  uint64_t seed = 1234 + 77 * k + cam;
  out.events.clear();
  const int n = 20000;
  for (int i = 0; i < n; ++i) {
    const uint32_t us = (uint32_t)((uint64_t)i * 33333 / n) + 33333u * k;
    const double tau = us * 1e-6;
    const double px = 100 + 120 * tau - (cam ? 6 : 0), py = 80 + 90 * tau;
    const uint64_t r = splitmix(seed);
    const int edge = r & 3;
    const double s = ((r >> 8) & 0xffff) / 65536.0 * 40.0;
    double ex = edge < 2 ? px + s : (edge == 2 ? px : px + 40);
    double ey = edge >= 2 ? py + s : (edge == 0 ? py : py + 40);
    dvs_msgs::Event e;
    e.x = (uint16_t)(ex + 0.5);
    e.y = (uint16_t)(ey + 0.5);
    e.ts.sec = 1700000000u + us / 1000000u;
    e.ts.nsec = (us % 1000000u) * 1000u;
    e.polarity = (edge == 1 || edge == 3) ? 1 : 0;
    out.events.push_back(e);
  }
}

int main() {
  esvio_fe_config cfg;
  esvio_fe_default_config(&cfg, 346, 260);
  cfg.max_events_per_window = 1 << 16;
  try {
    esvio::GpuFeatureTracker trackerData(cfg);
    dvs_msgs::EventArray L, R;
    size_t rows_total = 0;
    for (int k = 0; k < 6; ++k) {
      make_window(k, 0, L);
      make_window(k, 1, R);
      const dvs_msgs::Event& last = L.events.back();
      const double cur_time = (double)last.ts.sec + 1e-9 * (double)last.ts.nsec;  // node.cpp:190
      trackerData.PUB_THIS_FRAME = (k % 2 == 0);
      trackerData.trackEvent(cur_time, L, R);
      const auto rows = esvio::pack_feature_cloud(trackerData);
      std::map<int, std::vector<int>> seen;
      for (const auto& r : rows) {
        const int v = (int)(r.id_cam + 0.5f);  // stereo_estimator_node.cpp:388-401
        seen[v / 2].push_back(v % 2);
        if (r.z != 1.f) return 10;
      }
      for (const auto& kv : seen)
        if (kv.second[0] != 0 || kv.second.size() > 2 || (kv.second.size() == 2 && kv.second[1] != 1))
          return 11;  // feature_manager.cpp:331-340
      rows_total += rows.size();
      std::printf("window %d: %zu left, %zu right, %zu cloud rows\n", k, trackerData.ids.size(),
                  trackerData.ids_right.size(), rows.size());
    }
    if (trackerData.ids.empty() || rows_total == 0) return 12;
    // a group of two streams fed the same data must reproduce the single tracker, stream by stream
    esvio::GpuFeatureTracker single(cfg);
    esvio::GpuFeatureTrackerGroup group(cfg, 2);
    for (int k = 0; k < 4; ++k) {
      make_window(k, 0, L);
      make_window(k, 1, R);
      const dvs_msgs::Event& last = L.events.back();
      const double cur_time = (double)last.ts.sec + 1e-9 * (double)last.ts.nsec;
      single.PUB_THIS_FRAME = (k % 2 == 0);
      single.trackEvent(cur_time, L, R);
      group.tracker(0).PUB_THIS_FRAME = group.tracker(1).PUB_THIS_FRAME = (k % 2 == 0);
      group.trackEvents(std::vector<double>{cur_time, cur_time}, std::vector<dvs_msgs::EventArray>{L, L},
                        std::vector<dvs_msgs::EventArray>{R, R});
      for (int i = 0; i < 2; ++i) {
        const auto& t = group.tracker(i);
        if (t.ids != single.ids || t.ids_right != single.ids_right) return 13;
        for (size_t j = 0; j < t.cur_pts.size(); ++j)
          if (t.cur_pts[j].x != single.cur_pts[j].x || t.cur_pts[j].y != single.cur_pts[j].y) return 14;
      }
    }
    std::printf("group of 2 == single tracker\n");
    // frame path: FeatureTracker::trackImage through the same mirror (a tracker of its own, as
    // in stereo_image_tracker_node.cpp:45), on a blocky texture panning 2 px per frame
    struct Mat8 {  // the three members of cv::Mat the adapter touches
      uint8_t* data;
      size_t step;
      bool empty() const { return data == nullptr; }
    };
    esvio_fe_config icfg = cfg;
    icfg.max_cnt = 80, icfg.min_dist = 20;
    esvio::GpuFeatureTracker imageTracker(icfg);
    const int W = icfg.width, H = icfg.height, BW = W + 64;
    std::vector<uint8_t> big((size_t)BW * (H + 64));
    for (int y = 0; y < H + 64; ++y)
      for (int x = 0; x < BW; ++x) {
        uint32_t h = (uint32_t)(x / 12) * 2654435761u ^ (uint32_t)(y / 12) * 40503u;
        h ^= h >> 13, h *= 0x5bd1e995u, h ^= h >> 15;
        big[(size_t)y * BW + x] = (uint8_t)(40 + h % 180);
      }
    for (int k = 0; k < 3; ++k) {
      Mat8 left{&big[(size_t)(8 + 2 * k) * BW + 8 + 2 * k], (size_t)BW};
      Mat8 right{&big[(size_t)(8 + 2 * k) * BW + 12 + 2 * k], (size_t)BW};
      Mat8 none{nullptr, 0};
      imageTracker.PUB_THIS_FRAME = true;
      imageTracker.trackImage(0.05 * (k + 1), left, k == 1 ? none : right);
      if (k == 1 && !imageTracker.ids_right.empty()) return 16;
    }
    if (imageTracker.ids.empty()) return 15;
    {  // a Mat that knows its size (cv::Mat does) and is too small must be refused, not read
      struct MatRC {
        uint8_t* data;
        size_t step;
        int rows, cols;
        bool empty() const { return data == nullptr; }
      };
      MatRC small{big.data(), (size_t)BW, H - 1, W}, none{nullptr, 0, 0, 0};
      bool refused = false;
      try {
        imageTracker.trackImage(1.0, small, none);
      } catch (const std::invalid_argument&) {
        refused = true;
      }
      if (!refused) return 17;
    }
    std::printf("trackImage: %zu left / %zu right features\n", imageTracker.ids.size(),
                imageTracker.ids_right.size());
    return 0;
  } catch (const std::exception& e) {
    std::printf("%s\n", e.what());
    return std::strstr(e.what(), "no usable CUDA device") ? 3 : 1;
  }
}
