"""CPU oracle for the Sema vector-search hot path (test infrastructure only)."""
