"""py3 evaluation tooling around the hot path (the reference's script/*.py are Python 2 and need ROS)."""
