"""Stand-in for `einconv` (absent here): only `einconv.utils.get_conv_paddings`.  Test infrastructure."""
