package io.lubricant.consensus.raft.context.member;
/** Compile-time stand-in: Membership compares role CLASS TOKENS only (Follower.class). */
public class Follower implements io.lubricant.consensus.raft.RaftParticipant {}
