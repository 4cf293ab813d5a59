"""`cra5` import alias for the B200-native hot path (see dropin/README.md). Only the encode / decode surface exists."""
