// esvio_fe_adapter.hpp -- header-only C++ mirror of the reference's FeatureTracker (event path)
// on top of the C ABI in esvio_fe.h.  It keeps the member names the node reads after
// trackEvent (feature_tracker/src/feature_tracker.h:126-135; stereo_event_tracker_node.cpp:
// 289-323), so swapping `FeatureTracker trackerData` for `esvio::GpuFeatureTracker` in
// stereo_event_tracker_node.cpp:45 leaves handle_stereo_event and the PointCloud packing
// untouched.  No ROS / OpenCV / Eigen headers are needed: `EventArrayT` is any type with a
// contiguous `.events` container of 16-byte dvs_msgs::Event records
// (feature_tracker/src/dvs_msgs/Event.h:42-52), `Point2f` is layout-compatible with
// cv::Point2f.
#ifndef ESVIO_FE_ADAPTER_HPP
#define ESVIO_FE_ADAPTER_HPP

#include <stdexcept>
#include <string>
#include <vector>

#include "esvio_fe.h"

namespace esvio {

struct Point2f {
  float x, y;
};

class GpuFeatureTracker {
 public:
  // same role as readParameters_event + stereo_readIntrinsicParameter
  // (feature_tracker/src/parameters.cpp:183-282, feature_tracker.cpp:963-976)
  explicit GpuFeatureTracker(const esvio_fe_config& cfg) : cfg_(cfg) {
    const int rc = esvio_fe_create(&cfg_, &fe_);
    if (rc != ESVIO_FE_OK)
      throw std::runtime_error(std::string("esvio_fe_create: ") + esvio_fe_strerror(rc));
    const size_t m = (size_t)cfg_.max_cnt;
    id_.resize(m), cnt_.resize(m), idr_.resize(m);
    for (auto* v : {&u_, &v_, &unx_, &uny_, &vx_, &vy_, &ru_, &rv_, &runx_, &runy_, &rvx_, &rvy_})
      v->resize(m);
    out_.capacity = cfg_.max_cnt;
    out_.id = id_.data(), out_.track_cnt = cnt_.data();
    out_.u = u_.data(), out_.v = v_.data(), out_.un_x = unx_.data(), out_.un_y = uny_.data();
    out_.vx = vx_.data(), out_.vy = vy_.data();
    out_.id_right = idr_.data(), out_.ru = ru_.data(), out_.rv = rv_.data();
    out_.run_x = runx_.data(), out_.run_y = runy_.data(), out_.rvx = rvx_.data(),
    out_.rvy = rvy_.data();
  }
  ~GpuFeatureTracker() { esvio_fe_destroy(fe_); }
  GpuFeatureTracker(const GpuFeatureTracker&) = delete;
  GpuFeatureTracker& operator=(const GpuFeatureTracker&) = delete;

  // the reference's global PUB_THIS_FRAME (stereo_event_tracker_node.cpp:179,188)
  bool PUB_THIS_FRAME = true;

  // FeatureTracker::trackEvent(double, const EventArray&, const EventArray&)
  // (feature_tracker.h:51; call site stereo_event_tracker_node.cpp:193)
  template <class EventArrayT>
  void trackEvent(double _cur_time, const EventArrayT& event_left, const EventArrayT& event_right) {
    static_assert(sizeof(event_left.events[0]) == 16, "dvs_msgs::Event must be 16 bytes");
    esvio_events l{}, r{};
    l.aos = event_left.events.empty() ? nullptr : &event_left.events[0];
    l.n = event_left.events.size();
    r.aos = event_right.events.empty() ? nullptr : &event_right.events[0];
    r.n = event_right.events.size();
    const int rc = esvio_fe_track(fe_, _cur_time, &l, &r, PUB_THIS_FRAME ? 1 : 0, &out_);
    if (rc != ESVIO_FE_OK)
      throw std::runtime_error(std::string("esvio_fe_track: ") + esvio_fe_strerror(rc) + " (" +
                               esvio_fe_last_error(fe_) + ")");
    prev_time = cur_time;
    cur_time = _cur_time;
    unpack();
  }

  // FeatureTracker::trackEvent(double, const EventArray&, const EventArray&,
  // const Motion_correction_value) (feature_tracker.h:52; call site
  // stereo_event_tracker_node.cpp:254); the handle must have do_motion_correction = 1
  template <class EventArrayT>
  void trackEvent(double _cur_time, const EventArrayT& event_left, const EventArrayT& event_right,
                  const esvio_motion& measurements) {
    static_assert(sizeof(event_left.events[0]) == 16, "dvs_msgs::Event must be 16 bytes");
    esvio_events l{}, r{};
    l.aos = event_left.events.empty() ? nullptr : &event_left.events[0];
    l.n = event_left.events.size();
    r.aos = event_right.events.empty() ? nullptr : &event_right.events[0];
    r.n = event_right.events.size();
    const int rc = esvio_fe_track_mc(fe_, _cur_time, &l, &r, PUB_THIS_FRAME ? 1 : 0, &measurements, &out_);
    if (rc != ESVIO_FE_OK)
      throw std::runtime_error(std::string("esvio_fe_track_mc: ") + esvio_fe_strerror(rc) + " (" +
                               esvio_fe_last_error(fe_) + ")");
    prev_time = cur_time;
    cur_time = _cur_time;
    unpack();
  }

  // FeatureTracker::trackImage(double, const cv::Mat&, const cv::Mat&) (feature_tracker.h:49;
  // call site stereo_image_tracker_node.cpp:99).  MatT is cv::Mat or anything with
  // data / step / empty() of a CV_8UC1 image of the configured size; an empty right image is
  // the mono case (feature_tracker.cpp:245).  Use a tracker object of its own for frames, as
  // the reference's image node does.
  template <class MatT>
  void trackImage(double _cur_time, const MatT& img_left, const MatT& img_right) {
    // a frame smaller than the configured size would be an out-of-bounds host read in the copy
    if (!dims_ok(img_left, 0) || (!img_right.empty() && !dims_ok(img_right, 0)))
      throw std::invalid_argument("trackImage: frame size differs from the configured width x height");
    const uint8_t* r = img_right.empty() ? nullptr : (const uint8_t*)img_right.data;
    const int rc = esvio_fe_track_image(fe_, _cur_time, (const uint8_t*)img_left.data,
                                        (size_t)img_left.step, r, r ? (size_t)img_right.step : 0,
                                        PUB_THIS_FRAME ? 1 : 0, &out_);
    if (rc != ESVIO_FE_OK)
      throw std::runtime_error(std::string("esvio_fe_track_image: ") + esvio_fe_strerror(rc) +
                               " (" + esvio_fe_last_error(fe_) + ")");
    prev_time = cur_time;
    cur_time = _cur_time;
    unpack();
  }

  void reset() { esvio_fe_reset(fe_); }

 private:
  // rows / cols are checked when MatT has them (cv::Mat does); a bare {data, step} view is trusted
  template <class MatT>
  auto dims_ok(const MatT& m, int) const -> decltype((void)m.rows, (void)m.cols, bool()) {
    return m.rows == cfg_.height && m.cols == cfg_.width;
  }
  template <class MatT>
  bool dims_ok(const MatT&, long) const { return true; }

 public:

  // FeatureTracker::gettimesurface() (feature_tracker.cpp:894-897): CV_8U, row-major W x H
  std::vector<uint8_t> gettimesurface(int cam = 0) {
    std::vector<uint8_t> img((size_t)cfg_.width * cfg_.height);
    esvio_fe_time_surface(fe_, cam, img.data(), (size_t)cfg_.width);
    return img;
  }

  // public result members, names as in feature_tracker.h:126-135
  std::vector<int> ids, track_cnt, ids_right;
  std::vector<Point2f> cur_pts, cur_un_pts, pts_velocity;
  std::vector<Point2f> cur_right_pts, cur_un_right_pts, right_pts_velocity;
  double cur_time = 0.0, prev_time = 0.0;
  esvio_stats stats{};

  esvio_fe* handle() { return fe_; }

 private:
  void unpack() {
    const int nl = out_.n_left, nr = out_.n_right;
    ids.assign(id_.begin(), id_.begin() + nl);
    track_cnt.assign(cnt_.begin(), cnt_.begin() + nl);
    ids_right.assign(idr_.begin(), idr_.begin() + nr);
    cur_pts.resize(nl), cur_un_pts.resize(nl), pts_velocity.resize(nl);
    for (int i = 0; i < nl; ++i) {
      cur_pts[i] = {u_[i], v_[i]};
      cur_un_pts[i] = {unx_[i], uny_[i]};
      pts_velocity[i] = {vx_[i], vy_[i]};
    }
    cur_right_pts.resize(nr), cur_un_right_pts.resize(nr), right_pts_velocity.resize(nr);
    for (int i = 0; i < nr; ++i) {
      cur_right_pts[i] = {ru_[i], rv_[i]};
      cur_un_right_pts[i] = {runx_[i], runy_[i]};
      right_pts_velocity[i] = {rvx_[i], rvy_[i]};
    }
    stats = out_.stats;
  }

  esvio_fe_config cfg_;
  esvio_fe* fe_ = nullptr;
  esvio_tracks out_{};
  std::vector<int32_t> id_, cnt_, idr_;
  std::vector<float> u_, v_, unx_, uny_, vx_, vy_, ru_, rv_, runx_, runy_, rvx_, rvy_;
};

// Several independent stereo streams on one GPU (esvio_fe_group_*): the event stage of all
// streams shares its launches, results are those of separate trackers.  `tracker(i)` exposes the
// reference's member names for stream i after trackEvents().
class GpuFeatureTrackerGroup {
 public:
  struct Stream {
    bool PUB_THIS_FRAME = true;
    std::vector<int> ids, track_cnt, ids_right;
    std::vector<Point2f> cur_pts, cur_un_pts, pts_velocity;
    std::vector<Point2f> cur_right_pts, cur_un_right_pts, right_pts_velocity;
    esvio_tracks out{};
    std::vector<int32_t> id_, cnt_, idr_;
    std::vector<float> f_[12];
  };
  GpuFeatureTrackerGroup(const esvio_fe_config& cfg, int n_streams) : streams_(n_streams) {
    const int rc = esvio_fe_group_create(&cfg, n_streams, &g_);
    if (rc != ESVIO_FE_OK)
      throw std::runtime_error(std::string("esvio_fe_group_create: ") + esvio_fe_strerror(rc));
    const size_t m = (size_t)cfg.max_cnt;
    for (auto& s : streams_) {
      s.id_.resize(m), s.cnt_.resize(m), s.idr_.resize(m);
      for (auto& v : s.f_) v.resize(m);
      esvio_tracks& o = s.out;
      o.capacity = cfg.max_cnt;
      o.id = s.id_.data(), o.track_cnt = s.cnt_.data(), o.id_right = s.idr_.data();
      float** dst[12] = {&o.u, &o.v, &o.un_x, &o.un_y, &o.vx, &o.vy, &o.ru, &o.rv, &o.run_x, &o.run_y, &o.rvx, &o.rvy};
      for (int k = 0; k < 12; ++k) *dst[k] = s.f_[k].data();
    }
  }
  ~GpuFeatureTrackerGroup() { esvio_fe_group_destroy(g_); }
  GpuFeatureTrackerGroup(const GpuFeatureTrackerGroup&) = delete;
  GpuFeatureTrackerGroup& operator=(const GpuFeatureTrackerGroup&) = delete;

  Stream& tracker(int i) { return streams_[(size_t)i]; }
  int size() const { return (int)streams_.size(); }

  // one window of every stream: cur_time[i], left[i], right[i] as for trackEvent
  template <class EventArrayT>
  void trackEvents(const std::vector<double>& cur_time, const std::vector<EventArrayT>& left,
                   const std::vector<EventArrayT>& right) {
    const size_t S = streams_.size();
    std::vector<esvio_events> l(S), r(S);
    std::vector<int32_t> pub(S);
    std::vector<esvio_tracks> out(S);
    for (size_t i = 0; i < S; ++i) {
      l[i] = esvio_events{}, r[i] = esvio_events{};
      l[i].aos = left[i].events.empty() ? nullptr : &left[i].events[0];
      l[i].n = left[i].events.size();
      r[i].aos = right[i].events.empty() ? nullptr : &right[i].events[0];
      r[i].n = right[i].events.size();
      pub[i] = streams_[i].PUB_THIS_FRAME ? 1 : 0;
      out[i] = streams_[i].out;
    }
    const int rc = esvio_fe_group_track(g_, cur_time.data(), l.data(), r.data(), pub.data(), out.data());
    if (rc != ESVIO_FE_OK)
      throw std::runtime_error(std::string("esvio_fe_group_track: ") + esvio_fe_strerror(rc));
    for (size_t i = 0; i < S; ++i) {
      Stream& s = streams_[i];
      s.out = out[i];
      const int nl = s.out.n_left, nr = s.out.n_right;
      s.ids.assign(s.id_.begin(), s.id_.begin() + nl);
      s.track_cnt.assign(s.cnt_.begin(), s.cnt_.begin() + nl);
      s.ids_right.assign(s.idr_.begin(), s.idr_.begin() + nr);
      s.cur_pts.resize(nl), s.cur_un_pts.resize(nl), s.pts_velocity.resize(nl);
      for (int k = 0; k < nl; ++k) {
        s.cur_pts[k] = {s.f_[0][k], s.f_[1][k]};
        s.cur_un_pts[k] = {s.f_[2][k], s.f_[3][k]};
        s.pts_velocity[k] = {s.f_[4][k], s.f_[5][k]};
      }
      s.cur_right_pts.resize(nr), s.cur_un_right_pts.resize(nr), s.right_pts_velocity.resize(nr);
      for (int k = 0; k < nr; ++k) {
        s.cur_right_pts[k] = {s.f_[6][k], s.f_[7][k]};
        s.cur_un_right_pts[k] = {s.f_[8][k], s.f_[9][k]};
        s.right_pts_velocity[k] = {s.f_[10][k], s.f_[11][k]};
      }
    }
  }

 private:
  esvio_fe_group* g_ = nullptr;
  std::vector<Stream> streams_;
};

// One row of the `feature` sensor_msgs/PointCloud exactly as the node packs it
// (stereo_event_tracker_node.cpp:268-329): points[i] = (x, y, 1); channels = {id*2+cam, u, v,
// vx, vy}.  Left rows with track_cnt > 1 first, then right rows whose id was published left.
struct FeatureRow {
  float x, y, z, id_cam, u, v, vx, vy;
};

inline std::vector<FeatureRow> pack_feature_cloud(const GpuFeatureTracker& t) {
  std::vector<FeatureRow> rows;
  std::vector<int> pub;
  for (size_t j = 0; j < t.ids.size(); ++j)
    if (t.track_cnt[j] > 1) {
      pub.push_back(t.ids[j]);
      rows.push_back({t.cur_un_pts[j].x, t.cur_un_pts[j].y, 1.f, (float)(t.ids[j] * 2 + 0),
                      t.cur_pts[j].x, t.cur_pts[j].y, t.pts_velocity[j].x, t.pts_velocity[j].y});
    }
  for (size_t j = 0; j < t.ids_right.size(); ++j) {
    bool found = false;
    for (int id : pub) found |= (id == t.ids_right[j]);
    if (found)
      rows.push_back({t.cur_un_right_pts[j].x, t.cur_un_right_pts[j].y, 1.f,
                      (float)(t.ids_right[j] * 2 + 1), t.cur_right_pts[j].x, t.cur_right_pts[j].y,
                      t.right_pts_velocity[j].x, t.right_pts_velocity[j].y});
  }
  return rows;
}

}  // namespace esvio
#endif
