// esvio_fe_config.hpp -- loads the reference's configuration sets without OpenCV: what
// readParameters_event (feature_tracker/src/parameters.cpp:183-282) and
// FeatureTracker::stereo_readIntrinsicParameter (feature_tracker.cpp:963-976) read at node
// start-up -- config/<set>/es*io.yaml plus the two camodocal PINHOLE calibrations it names --
// into an esvio_fe_config.  Header-only, C++11; the Python twin is esvio_b200/config.py.
//
// A node that still links OpenCV can keep its own cv::FileStorage code and copy the globals
// (INTEGRATION.md); this header is for the ROS-free node (esvio_fe_node.hpp) and for tools.
// The reader covers the OpenCV FileStorage YAML 1.0 dialect of those files: `%YAML:1.0`,
// nested maps by indentation, `!!opencv-matrix` maps, flow sequences that may span lines,
// `#` comments.  Like cv::FileNode, a missing numeric key reads as 0.
#pragma once
#include <cstdlib>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "esvio_fe.h"

namespace esvio {

// flat view of an OpenCV-YAML file: "key" or "map.key" -> raw scalar text; flow sequences
// ("data: [ ... ]") are kept as their comma-separated text
class YamlDoc {
 public:
  explicit YamlDoc(const std::string& path) {
    std::ifstream f(path.c_str());
    if (!f) throw std::runtime_error("cannot open " + path);
    std::vector<std::pair<int, std::string>> stack;  // (indent, prefix)
    std::string line, pending_key, pending;
    while (std::getline(f, line)) {
      line = strip_comment(line);
      if (!pending_key.empty()) {
        pending += " " + trim(line);
        if (line.find(']') != std::string::npos) {
          kv_[pending_key] = inside_brackets(pending);
          pending_key.clear();
        }
        continue;
      }
      const std::string body = trim(line);
      if (body.empty() || body[0] == '%' || body == "---") continue;
      const size_t colon = body.find(':');
      if (colon == std::string::npos) continue;
      const int indent = (int)line.find_first_not_of(" \t");
      while (!stack.empty() && indent <= stack.back().first) stack.pop_back();
      const std::string key = (stack.empty() ? std::string() : stack.back().second + ".") + trim(body.substr(0, colon));
      const std::string val = trim(body.substr(colon + 1));
      if (val.empty() || val.compare(0, 2, "!!") == 0) {
        stack.push_back(std::make_pair(indent, key));
      } else if (val[0] == '[') {
        if (val.find(']') != std::string::npos) kv_[key] = inside_brackets(val);
        else pending_key = key, pending = val;
      } else {
        kv_[key] = unquote(val);
      }
    }
  }
  bool has(const std::string& k) const { return kv_.count(k) != 0; }
  std::string str(const std::string& k) const {
    std::map<std::string, std::string>::const_iterator it = kv_.find(k);
    return it == kv_.end() ? std::string() : it->second;
  }
  double num(const std::string& k) const {  // missing or non-numeric: 0, like cv::FileNode
    const std::string s = str(k);
    if (s.empty()) return 0.0;
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    return end == s.c_str() ? 0.0 : v;
  }
  std::vector<double> seq(const std::string& k) const {
    std::vector<double> out;
    const std::string s = str(k);
    size_t pos = 0;
    while (pos < s.size()) {
      size_t comma = s.find(',', pos);
      if (comma == std::string::npos) comma = s.size();
      const std::string tok = trim(s.substr(pos, comma - pos));
      if (!tok.empty()) out.push_back(std::strtod(tok.c_str(), nullptr));
      pos = comma + 1;
    }
    return out;
  }

 private:
  static std::string trim(const std::string& s) {
    const size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
  }
  static std::string unquote(const std::string& s) {
    if (s.size() >= 2 && s[0] == s[s.size() - 1] && (s[0] == '"' || s[0] == '\'')) return s.substr(1, s.size() - 2);
    return s;
  }
  static std::string strip_comment(const std::string& s) {
    char quote = 0;
    for (size_t i = 0; i < s.size(); ++i) {
      if (quote) {
        if (s[i] == quote) quote = 0;
      } else if (s[i] == '"' || s[i] == '\'') {
        quote = s[i];
      } else if (s[i] == '#') {
        return s.substr(0, i);
      }
    }
    return s;
  }
  static std::string inside_brackets(const std::string& s) {
    const size_t a = s.find('['), b = s.rfind(']');
    return s.substr(a + 1, b - a - 1);
  }
  std::map<std::string, std::string> kv_;
};

// camodocal PINHOLE calibration (PinholeCamera::Parameters::readFromYamlFile,
// camera_model/src/camera_models/PinholeCamera.cc:150-191)
inline void read_pinhole_yaml(const std::string& path, esvio_pinhole* cam) {
  const YamlDoc y(path);
  if (y.has("model_type") && y.str("model_type") != "PINHOLE")
    throw std::runtime_error(path + ": model_type " + y.str("model_type") + " (the event front-end lifts with PINHOLE only)");
  cam->k1 = y.num("distortion_parameters.k1");
  cam->k2 = y.num("distortion_parameters.k2");
  cam->p1 = y.num("distortion_parameters.p1");
  cam->p2 = y.num("distortion_parameters.p2");
  cam->fx = y.num("projection_parameters.fx");
  cam->fy = y.num("projection_parameters.fy");
  cam->cx = y.num("projection_parameters.cx");
  cam->cy = y.num("projection_parameters.cy");
}

struct EventNodeParams {  // what the node itself keeps from the file
  int freq, show_track, max_cnt_img, min_dist_img, image_width, image_height;
  std::string event_left_topic, event_right_topic, imu_topic;
};

// readParameters_event + stereo_readIntrinsicParameter.  `esvio_folder`: the ROS param of the
// same name that prefixes the calibration files (parameters.cpp:192,243-244); empty = the
// directory of `config_file`.
inline void read_parameters_event(const std::string& config_file, const std::string& esvio_folder,
                                  esvio_fe_config* cfg, EventNodeParams* node = nullptr) {
  const YamlDoc y(config_file);
  std::string folder = esvio_folder;
  if (folder.empty()) {
    const size_t slash = config_file.find_last_of('/');
    folder = slash == std::string::npos ? std::string(".") : config_file.substr(0, slash);
  }
  esvio_fe_default_config(cfg, (int32_t)y.num("event_width"), (int32_t)y.num("event_height"));
  cfg->max_cnt = (int32_t)y.num("max_cnt");
  cfg->min_dist = (int32_t)y.num("min_dist");
  cfg->flow_back = (int32_t)y.num("flow_back");
  cfg->equalize = (int32_t)y.num("equalize");
  cfg->f_threshold = y.num("F_threshold");
  cfg->ts_lk_threshold = y.num("TS_LK_threshold");
  cfg->decay_ms = y.num("decay_ms");
  cfg->ignore_polarity = (int32_t)y.num("ignore_polarity");
  cfg->median_blur_kernel_size = (int32_t)y.num("median_blur_kernel_size");
  cfg->feature_filter_threshold = y.num("feature_filter_threshold");
  cfg->do_motion_correction = (int32_t)y.num("Do_motion_correction");
  cfg->focal_length = 460.0;  // FOCAL_LENGTH = 460 (parameters.cpp:274)
  read_pinhole_yaml(folder + "/" + y.str("event_left_calib"), &cfg->cam[0]);
  read_pinhole_yaml(folder + "/" + y.str("event_right_calib"), &cfg->cam[1]);
  if (node) {
    node->freq = (int)y.num("freq");
    if (node->freq == 0) node->freq = 100;  // parameters.cpp:277-278
    node->show_track = (int)y.num("show_track");
    node->max_cnt_img = (int)y.num("max_cnt_img");
    node->min_dist_img = (int)y.num("min_dist_img");
    node->image_width = (int)y.num("image_width");
    node->image_height = (int)y.num("image_height");
    node->event_left_topic = y.str("event_left_topic");
    node->event_right_topic = y.str("event_right_topic");
    node->imu_topic = y.str("imu_topic");
  }
}

}  // namespace esvio
