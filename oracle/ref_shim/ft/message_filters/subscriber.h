// Stand-in: message_filters is included by the node and not used.
#pragma once
