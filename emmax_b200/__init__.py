"""emmax_b200 — B200-native `generate_actions` hot path of Emma-X behind the reference's Python surface."""
