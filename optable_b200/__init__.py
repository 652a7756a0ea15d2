"""optable_b200: B200-native ray-propagation back end behind optable's Python API."""
