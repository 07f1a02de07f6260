// ref_windower_api.cc -- TEST INFRASTRUCTURE: the reference's own event re-packing tool,
// dependences/events_repacking_helper/src/EventMessageEditor.cpp, included UNMODIFIED from where it
// lies (its main() renamed; the struct EventMessageEditor is defined in that .cpp, so it has to be
// in this translation unit) and driven event by event.  rosbag is a stand-in whose Bag::write hands
// each written EventArray to this file.  Built into oracle/_ref/libesvio_ref_node.so; used only by
// tests/test_oracle_ref_node.py to pin esvio_b200/node.py's EventWindower.
#define main esvio_ref_eme_main
#include "EventMessageEditor.cpp"
#undef main

#include <vector>

struct Written {
  uint32_t stamp_sec, stamp_nsec, hdr_sec, hdr_nsec, n;
  uint32_t first_sec, first_nsec, last_sec, last_nsec;
};
struct Windower {
  std::string topic = "/events";
  EventMessageEditor editor;
  rosbag::Bag bag;
  std::vector<Written> out;
  explicit Windower(double frequency) : editor(frequency, topic) {}
};
static std::vector<Windower*> g_windowers;

void esvio_ref_shim_bag_write(void* bag, const char*, const ros::Time& stamp, const dvs_msgs::EventArray& msg) {
  for (Windower* w : g_windowers)
    if (&w->bag == bag) {
      Written r{};
      r.stamp_sec = stamp.sec, r.stamp_nsec = stamp.nsec;
      r.hdr_sec = msg.header.stamp.sec, r.hdr_nsec = msg.header.stamp.nsec;
      r.n = (uint32_t)msg.events.size();
      if (r.n) {
        r.first_sec = msg.events.front().ts.sec, r.first_nsec = msg.events.front().ts.nsec;
        r.last_sec = msg.events.back().ts.sec, r.last_nsec = msg.events.back().ts.nsec;
      }
      w->out.push_back(r);
    }
}

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void* ref_eme_create(double frequency) {
  Windower* w = new Windower(frequency);
  g_windowers.push_back(w);
  return w;
}
REF_API void ref_eme_destroy(void* p) {
  Windower* w = static_cast<Windower*>(p);
  for (size_t i = 0; i < g_windowers.size(); ++i)
    if (g_windowers[i] == w) g_windowers.erase(g_windowers.begin() + i);
  delete w;
}
// EventMessageEditor::insertEvent for every event of a chunk (EventMessageEditor.cpp:34-50,117-120)
REF_API void ref_eme_insert(void* p, const uint16_t* x, const uint16_t* y, const uint32_t* sec, const uint32_t* nsec,
                            const uint8_t* pol, size_t n) {
  Windower* w = static_cast<Windower*>(p);
  for (size_t i = 0; i < n; ++i) {
    dvs_msgs::Event e;
    e.x = x[i], e.y = y[i], e.ts.sec = sec[i], e.ts.nsec = nsec[i], e.polarity = pol[i];
    w->editor.insertEvent(e, &w->bag);
  }
}
REF_API int ref_eme_count(void* p) { return (int)static_cast<Windower*>(p)->out.size(); }
// 9 words per written message: write stamp (sec, nsec), header stamp (sec, nsec), events, first / last event time
REF_API void ref_eme_get(void* p, int i, uint32_t* out9) {
  const Written& r = static_cast<Windower*>(p)->out[i];
  const uint32_t v[9] = {r.stamp_sec, r.stamp_nsec, r.hdr_sec, r.hdr_nsec, r.n, r.first_sec, r.first_nsec, r.last_sec, r.last_nsec};
  for (int k = 0; k < 9; ++k) out9[k] = v[k];
}
