"""alan_b200 -- B200-native engine for alan's logPQ plate-tree reduction."""
