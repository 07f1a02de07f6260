// Prints what include/esvio_fe_config.hpp reads from a configuration set, one "key value" per
// line with 17 significant digits (tests/test_config.py compares it with esvio_b200/config.py).
// esvio_fe_default_config is the library's; this tool must not need the GPU library, so a
// stand-in with the same defaults is linked instead.
#include <cstdio>
#include <cstring>

#include "esvio_fe_config.hpp"

extern "C" void esvio_fe_default_config(esvio_fe_config* c, int32_t width, int32_t height) {
  std::memset(c, 0, sizeof(*c));
  c->width = width;
  c->height = height;
  c->focal_length = 460.0;
  c->use_ransac = 1;
  c->max_events_per_window = 1 << 20;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  esvio_fe_config c;
  esvio::EventNodeParams n;
  try {
    esvio::read_parameters_event(argv[1], argc > 2 ? argv[2] : "", &c, &n);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  std::printf("width %d\nheight %d\nmax_cnt %d\nmin_dist %d\nflow_back %d\nequalize %d\n", c.width, c.height,
              c.max_cnt, c.min_dist, c.flow_back, c.equalize);
  std::printf("f_threshold %.17g\nts_lk_threshold %.17g\ndecay_ms %.17g\nignore_polarity %d\n", c.f_threshold,
              c.ts_lk_threshold, c.decay_ms, c.ignore_polarity);
  std::printf("median_blur_kernel_size %d\nfeature_filter_threshold %.17g\ndo_motion_correction %d\n",
              c.median_blur_kernel_size, c.feature_filter_threshold, c.do_motion_correction);
  std::printf("focal_length %.17g\nfreq %d\nmax_cnt_img %d\nmin_dist_img %d\n", c.focal_length, n.freq,
              n.max_cnt_img, n.min_dist_img);
  for (int i = 0; i < 2; ++i)
    std::printf("cam%d %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", i, c.cam[i].fx, c.cam[i].fy,
                c.cam[i].cx, c.cam[i].cy, c.cam[i].k1, c.cam[i].k2, c.cam[i].p1, c.cam[i].p2);
  std::printf("event_left_topic %s\n", n.event_left_topic.c_str());
  return 0;
}
