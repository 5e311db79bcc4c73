// Stand-in for <rclcpp/rclcpp.hpp>: nothing of it is used by the files compiled into oracle/_ref.
