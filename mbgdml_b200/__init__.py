"""mbgdml_b200: B200-native many-body GDML prediction path."""
