// Stand-in for std_msgs/Int32 (included by EventMessageEditor.cpp, unused).
#pragma once
