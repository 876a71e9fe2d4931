package io.lubricant.consensus.raft.context.member;
/** Compile-time stand-in: Membership compares role CLASS TOKENS only (Candidate.class). */
public class Candidate implements io.lubricant.consensus.raft.RaftParticipant {}
