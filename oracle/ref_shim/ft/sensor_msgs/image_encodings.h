// Stand-in: nothing of sensor_msgs/image_encodings.h is used on the tracking path.
#pragma once
