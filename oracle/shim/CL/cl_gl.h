/* empty: shim for the reference host sources (oracle only) */
