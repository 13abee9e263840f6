"""cim_b200 -- B200-native CIM proposal-level hot path."""
