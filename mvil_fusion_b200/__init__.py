"""B200-native hot path of a sliding-window visual-inertial-LiDAR estimator."""
