"""Mirror of openpoints/AMContrast3D/AEF (ambiguity estimation framework)."""
