// Stand-in for std_msgs/Bool.  TEST INFRASTRUCTURE.
#pragma once
namespace std_msgs {
struct Bool {
  bool data = false;
};
}  // namespace std_msgs
