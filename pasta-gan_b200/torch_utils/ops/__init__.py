"""B200-native operator layer with the reference's ``torch_utils.ops`` call surface."""
