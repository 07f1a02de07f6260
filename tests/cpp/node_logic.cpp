// Host-logic check of include/esvio_fe_node.hpp (no GPU, no ROS): a scripted pair of raw event
// streams goes through EventWindower -> EventPairer -> StereoEventNode with a stand-in tracker,
// and every decision is printed.  tests/test_node_logic.py replays the same script through the
// Python twin (esvio_b200/node.py) and compares the traces line by line.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "esvio_fe_node.hpp"

namespace dvs_msgs {
struct Time {
  uint32_t sec, nsec;
};
struct Event {
  uint16_t x, y;
  Time ts;
  uint8_t polarity;
};
struct EventArray {
  double stamp = 0.0;
  std::vector<Event> events;
};
}  // namespace dvs_msgs

struct P2 {
  float x, y;
};
struct StubTracker {
  bool PUB_THIS_FRAME = false;
  std::vector<int> ids, track_cnt, ids_right;
  std::vector<P2> cur_pts, cur_un_pts, pts_velocity, cur_right_pts, cur_un_right_pts, right_pts_velocity;
  int calls = 0;
  void fill(double t, size_t nl, size_t nr) {
    ++calls;
    const int n = 3 + calls % 4;
    ids.clear(), track_cnt.clear(), ids_right.clear();
    cur_pts.clear(), cur_un_pts.clear(), pts_velocity.clear();
    cur_right_pts.clear(), cur_un_right_pts.clear(), right_pts_velocity.clear();
    for (int i = 0; i < n; ++i) {
      ids.push_back(calls + i);
      track_cnt.push_back(1 + (i + calls) % 3);
      cur_pts.push_back({(float)i, (float)(nl % 100)});
      cur_un_pts.push_back({0.1f * i, 0.2f * i});
      pts_velocity.push_back({1.f, 2.f});
      if (i % 2 == 0) {
        ids_right.push_back(calls + i);
        cur_right_pts.push_back({(float)i - 5.f, (float)(nr % 100)});
        cur_un_right_pts.push_back({0.1f * i - 0.05f, 0.2f * i});
        right_pts_velocity.push_back({1.5f, 2.5f});
      }
    }
    std::printf("T %.9f nl=%zu nr=%zu pub=%d\n", t, nl, nr, PUB_THIS_FRAME ? 1 : 0);
  }
  template <class ImgT>
  void trackImage(double t, const ImgT& l, const ImgT& r) {
    fill(t, (size_t)l.tag, (size_t)r.tag);
  }
  void trackEvent(double t, const dvs_msgs::EventArray& l, const dvs_msgs::EventArray& r) {
    fill(t, l.events.size(), r.events.size());
  }
  void trackEvent(double t, const dvs_msgs::EventArray& l, const dvs_msgs::EventArray& r,
                  const esvio_motion& m) {
    std::printf("M a=%.6f,%.6f,%.6f w=%.6f,%.6f,%.6f v=%.6f vp=%.6f t1=%.9f\n", m.accel[0], m.accel[1],
                m.accel[2], m.omega[0], m.omega[1], m.omega[2], m.state_v[0], m.v_pre[0], m.t1);
    fill(t, l.events.size(), r.events.size());
  }
};

// scripted raw stream: `n` events from t0, one every `step_us` microseconds, with a gap
static std::vector<dvs_msgs::Event> make_stream(uint32_t sec0, uint32_t us0, int n, int step_us,
                                                int gap_at, int gap_us) {
  std::vector<dvs_msgs::Event> v;
  uint64_t us = (uint64_t)us0;
  for (int i = 0; i < n; ++i) {
    if (i == gap_at) us += (uint64_t)gap_us;
    dvs_msgs::Event e;
    e.x = (uint16_t)((i * 7) % 346);
    e.y = (uint16_t)((i * 13) % 260);
    e.ts.sec = sec0 + (uint32_t)(us / 1000000u);
    e.ts.nsec = (uint32_t)(us % 1000000u) * 1000u;
    e.polarity = (uint8_t)(i & 1);
    v.push_back(e);
    us += (uint64_t)step_us;
  }
  return v;
}

// stand-in for one sensor_msgs/Image after getImageFromMsg: a stamp and a tag
struct ImageMsg {
  double stamp = 0.0;
  int tag = 0;
};

// image node script: 20 Hz stereo frames; right frame 7 is lost; left frame 30 arrives exactly
// 1 s older than the waiting right frame (thrown by the `<=` rule); a 1.5 s hole after frame 40
// restarts the node; the last frame goes back in time (second restart)
static int image_script() {
  StubTracker trk;
  esvio::StereoImageNode<StubTracker> node(trk, 10);
  esvio::ImagePairer<ImageMsg> pairer;
  const double t0 = 1700000000.0;
  auto step = [&](bool left, double stamp, int tag) {
    ImageMsg m;
    m.stamp = stamp, m.tag = tag;
    if (left) pairer.pushLeft(m);
    else pairer.pushRight(m);
    while (pairer.ready()) {
      ImageMsg l, r;
      double ts;
      if (!pairer.poll(&l, &r, &ts)) {
        std::printf("D\n");
        continue;
      }
      esvio::FeatureCloud cloud;
      const bool pub = node.handle_stereo_image(l, r, ts, &cloud);
      std::printf("H %.9f l=%d r=%d published=%d rows=%zu restarts=%d", ts, l.tag, r.tag, pub ? 1 : 0,
                  pub ? cloud.rows.size() : 0, node.restarts);
      if (pub)
        for (const auto& row : cloud.rows) std::printf(" %g:%g", row.id_cam, row.u);
      std::printf("\n");
    }
  };
  for (int k = 0; k < 60; ++k) {
    double tl = t0 + 0.05 * k + (k > 40 ? 1.5 : 0.0);
    double tr = tl + 0.002;
    if (k == 30) tr = tl + 1.0;  // left exactly 1 s older than right: left is thrown
    if (k == 59) tl = tr = t0 + 0.05 * 50;  // back in time
    step(true, tl, 1000 + k);
    if (k != 7) step(false, tr, 2000 + k);
  }
  std::printf("END dropped=%d restarts=%d tracked=%d\n", pairer.dropped, node.restarts,
              node.frames_tracked);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'i') return image_script();
  const bool mc = argc > 1 && argv[1][0] == 'm';
  using EA = dvs_msgs::EventArray;
  std::vector<EA> lm, rm;
  {
    esvio::EventWindower<EA> wl(30.0), wr(30.0);
    // left: 60000 events, 25 us apart (1.5 s), a 1.4 s hole after event 30000
    for (const auto& e : make_stream(1700000000u, 100, 60000, 25, 30000, 1400000))
      wl.insertEvent(e, [&](EA&& m) { lm.push_back(std::move(m)); });
    // right: starts 3 ms later, same rate, same hole, 30 events fewer
    for (const auto& e : make_stream(1700000000u, 3100, 59970, 25, 30000, 1400000))
      wr.insertEvent(e, [&](EA&& m) { rm.push_back(std::move(m)); });
  }
  for (const auto& m : lm) std::printf("WL %.9f %zu\n", m.stamp, m.events.size());
  for (const auto& m : rm) std::printf("WR %.9f %zu\n", m.stamp, m.events.size());

  StubTracker trk;
  esvio::StereoEventNode<StubTracker> node(trk, 15, mc);
  if (mc) {
    for (int i = 0; i < 400; ++i)
      node.motion.pushImu({1700000000.0 + 0.005 * i, 0.01 * i, -0.02 * i, 0.5});
    node.motion.pushImu({1700000000.0, 9, 9, 9});  // out of order: ignored
    for (int i = 0; i < 40; ++i)
      node.motion.pushOdometry({1700000000.0 + 0.05 * i, 0.1 * i * i, 0.2, -0.1 * i});
  }
  esvio::EventPairer<EA> pairer;
  size_t il = 0, ir = 0;
  while (il < lm.size() || ir < rm.size()) {
    const bool left = ir >= rm.size() || (il < lm.size() && lm[il].stamp <= rm[ir].stamp);
    if (left) pairer.pushLeft(lm[il++]);
    else pairer.pushRight(rm[ir++]);
    while (pairer.ready()) {
      EA l, r;
      double ts;
      if (!pairer.poll(&l, &r, &ts)) continue;
      esvio::FeatureCloud cloud;
      const bool pub = node.handle_stereo_event(l, r, ts, &cloud);
      std::printf("H %.9f published=%d rows=%zu", ts, pub ? 1 : 0, pub ? cloud.rows.size() : 0);
      if (pub)
        for (const auto& row : cloud.rows) std::printf(" %g:%g", row.id_cam, row.u);
      std::printf("\n");
    }
  }
  // a jump of more than a second resets the node; the next pair is a "first" one again
  {
    EA l = lm.back(), r = rm.back();
    const double base = l.stamp;
    const double stamps[4] = {base + 2.0, base + 2.033, base + 2.066, base + 2.0};
    for (double ts : stamps) {
      esvio::FeatureCloud cloud;
      const bool pub = node.handle_stereo_event(l, r, ts, &cloud);
      std::printf("H %.9f published=%d rows=%zu restarts=%d\n", ts, pub ? 1 : 0,
                  pub ? cloud.rows.size() : 0, node.restarts);
    }
    EA empty;
    esvio::FeatureCloud cloud;
    std::printf("E %d\n", node.handle_stereo_event(empty, r, base + 3.0, &cloud) ? 1 : 0);
  }
  std::printf("END dropped=%d restarts=%d tracked=%d\n", pairer.dropped, node.restarts,
              node.windows_tracked);
  return 0;
}
