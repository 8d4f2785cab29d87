"""The worked polynomial example of the reference (binf/example/), lowered to the device."""
